"""Test helper: load the reference's pickled PSM/spectrum fixtures (only where /root/reference exists).

The spectra pickles reference pyteomics.auxiliary.structures.unitfloat; pyteomics is not
installed, so a stub class is registered before unpickling.
"""
import os
import pickle
import sys
import types

import numpy as np

REF_ROOT = "/root/reference"
PAIR_DIR = os.path.join(REF_ROOT, "test", "match_spectra_pairs")


def _stub_pyteomics():
    if "pyteomics.auxiliary.structures" in sys.modules:
        return
    m = types.ModuleType("pyteomics")
    a = types.ModuleType("pyteomics.auxiliary")
    s = types.ModuleType("pyteomics.auxiliary.structures")

    class unitfloat(float):
        def __new__(cls, v, unit_info=None):
            o = float.__new__(cls, v)
            o.unit_info = unit_info
            return o

    s.unitfloat = unitfloat
    m.auxiliary = a
    a.structures = s
    sys.modules.update({"pyteomics": m, "pyteomics.auxiliary": a, "pyteomics.auxiliary.structures": s})


def have_reference_fixtures():
    return os.path.isdir(PAIR_DIR)


def load_pairs():
    """-> list of dicts(peptide, n_mod, charge, mz, inten, aux_pos, aux_mass, tag)"""
    _stub_pyteomics()
    out = []
    for tag in ["1_mods", "2_mods", "3_mods", "aux"]:
        with open(os.path.join(PAIR_DIR, "velos_matches_%s.pkl" % tag), "rb") as f:
            M = pickle.load(f)
        with open(os.path.join(PAIR_DIR, "velos_spectra_%s.pkl" % tag), "rb") as f:
            S = pickle.load(f)
        for mt, sp in zip(M, S):
            # every mod in these fixtures (also the "aux" file) is phospho 79.966331; the
            # reference test passes len(mod_positions) as n_of_mod (test/test_ascore.py:25-29)
            out.append(dict(tag=tag, peptide=mt["peptide"], charge=int(mt["charge_state"]),
                            n_mod=len(mt["mod_positions"]),
                            mz=np.asarray(sp["mz_values"], dtype=np.float64).copy(),
                            inten=np.asarray(sp["intensity_values"], dtype=np.float64).copy()))
    return out
