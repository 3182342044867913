"""Generate tests/golden/cli/config1.npz: BASELINE config 1 (the reference's example mzML + pepXML
through its CLI defaults) as data that travels to the GPU box.

Run in the build container (needs /root/reference and oracle/_ref):

    python tests/golden/make_config1.py

The file holds (a) the ten example spectra and the twenty pepXML hits as arrays / JSON, read
from the reference's example files with pyascore_b200.parsing, and (b) the TSV the reference
produces for several CLI settings.  (b) is computed by replaying the reference's own driver loop
(pyascore/__main__.py:129-164) around the compiled, unmodified reference scorer
(oracle.cscorer.RefPyAscore) -- `import pyascore` itself needs pyteomics, which is absent here.
"""
import json
import os
import sys
from itertools import groupby

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.cscorer import RefPyAscore  # noqa: E402
from pyascore_b200.parsing import COMMON_MODS, IdentificationParser, MassCorrector, SpectraParser  # noqa: E402
from pyascore_b200.parsing import _xml  # noqa: E402

EX = "/root/reference/test/example_inputs"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cli", "config1.npz")

SETTINGS = {
    "default": {},
    "hit_depth2": {"hit_depth": 2},
    "hires_nl": {"mz_error": 0.02, "neutral_loss_groups": "st", "neutral_loss_masses": "97.9769", "hit_depth": -1},
    "charge1_cz": {"max_fragment_charge": 1, "fragment_types": "cz"},
    "oxidation": {"residues": "M", "mod_mass": 15.9949, "hit_depth": 2},
}
DEFAULTS = dict(residues="STY", mod_mass=79.966331, mz_error=0.5, mod_correction_tol=1., zero_based=False,
                neutral_loss_groups="", neutral_loss_masses="", static_mod_groups="C", static_mod_masses="57.021464",
                fragment_types="by", max_fragment_charge=5, hit_depth=1)


def reference_rows(spectra_map, a):
    """the reference's driver loop, restated around the compiled reference scorer"""
    static = {}
    for g, m in zip(a["static_mod_groups"].split(","), a["static_mod_masses"].split(",")):
        static.update({aa: float(m) for aa in g})
    mods = COMMON_MODS.copy()
    mods.update({aa: a["mod_mass"] for aa in a["residues"]})
    mods.update(static)
    psms = sorted(IdentificationParser(EX + "/psms/test_psms.pep.xml", "pepXML", MassCorrector(mod_mass_dict=mods),
                                       static_mods=static).to_list(), key=lambda m: m["scan"])
    sc = RefPyAscore(100., 10, a["residues"], a["mod_mass"], a["mz_error"], a["fragment_types"])
    if a["neutral_loss_groups"] and a["neutral_loss_masses"]:
        for g, m in zip(a["neutral_loss_groups"].split(","), a["neutral_loss_masses"].split(",")):
            sc.add_neutral_loss(g, float(m))
    rows = []
    for _, group in groupby(psms, lambda x: x["scan"]):
        for ind, match in enumerate(group):
            if ind == a["hit_depth"]:
                break
            spec = spectra_map[match["scan"]]
            cpos, cmass, nvar = [], [], 0
            for pos, mass in zip(match["mod_positions"], match["mod_masses"]):
                shift = 1 if a["zero_based"] else 0
                aa = "n" if pos + shift == 0 else match["peptide"][pos - 1 + shift]
                if np.isclose(a["mod_mass"], mass, rtol=1e-6, atol=a["mod_correction_tol"]) and aa in a["residues"]:
                    nvar += 1
                else:
                    cpos.append(pos + shift); cmass.append(mass)
            z = match["charge_state"] or spec["precursor_charge"] or 2
            z = max(z, 2)
            if nvar > 0:
                sc.score(spec["mz_values"], spec["intensity_values"], match["peptide"], nvar,
                         min(a["max_fragment_charge"], z - 1), np.array(cpos, np.uint32), np.array(cmass, np.float32))
                rows.append("\t".join([str(match["scan"]), sc.best_sequence, repr(float(sc.best_score)),
                                       ";".join(str(s) for s in sc.ascores),
                                       ";".join(",".join(str(x) for x in sl) for sl in sc.alt_sites)]))
    return "Scan\tLocalizedSequence\tPepScore\tAscores\tAltSites\n" + "".join(r + "\n" for r in rows)


def main():
    spectra = SpectraParser(EX + "/spectra/test_spectra.mzML", "mzML").to_list()
    # the literals pinned by the reference's own test (test/test_spec_parsers.py:5-8)
    assert [s["scan"] for s in spectra] == [14760, 18330, 20462, 21996, 26219, 26962, 27845, 31328, 32257, 35669]
    spectra_map = {s["scan"]: s for s in spectra}
    queries = []
    for scan, charge, hits in _xml.iter_pepxml(EX + "/psms/test_psms.pep.xml", "xcorr_score"):
        queries.append(dict(scan=scan, charge=charge, hits=[
            dict(peptide=p, score=repr(float(sc)), mods=[[int(a), repr(float(np.float64(round(float(b), 2))))] for a, b in zip(pos, mass)])
            for p, sc, pos, mass in hits]))
    arrays = dict(
        scans=np.array([s["scan"] for s in spectra], np.int64),
        precursor_mz=np.array([s["precursor_mz"] for s in spectra], np.float64),
        precursor_charge=np.array([s["precursor_charge"] for s in spectra], np.int32),
        spec_off=np.concatenate([[0], np.cumsum([s["mz_values"].size for s in spectra])]).astype(np.int64),
        mz=np.concatenate([s["mz_values"] for s in spectra]),
        inten=np.concatenate([s["intensity_values"] for s in spectra]).astype(np.float32),   # 32-bit in the file
        queries=np.frombuffer(json.dumps(queries).encode(), np.uint8))
    expected = {}
    for name, over in SETTINGS.items():
        a = dict(DEFAULTS, **over)
        expected[name] = dict(args=over, tsv=reference_rows(spectra_map, a))
        print("== %s %s\n%s" % (name, over, expected[name]["tsv"]))
    arrays["expected"] = np.frombuffer(json.dumps(expected).encode(), np.uint8)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **arrays)
    print("wrote %s (%.0f KB)" % (OUT, os.path.getsize(OUT) / 1e3))


if __name__ == "__main__":
    main()
