"""Pin the percolatorTXT / mokapotTXT extractors to the REFERENCE's own code (pyascore/parsing/id_parsers.py:183-221,
:475-653, :656-815), run here on small tables:

    python tests/golden/make_idparse.py        (needs /root/reference; writes tests/golden/idparse/*.json)

The reference module is imported from /root/reference by path.  What it needs from the environment and this image
lacks is stubbed, nothing in its own code is touched:
  * pyteomics (not installed): `mass.std_aa_mass` is the 5-decimal residue table pyteomics ships (the XML readers
    `mzid.MzIdentML` / `pepxml.PepXML` are never reached by the two text formats);
  * pandas 3 drops the grouping column from the frames `groupby(col).apply(f)` hands to f (pandas < 2.2, which the
    reference was written for, keeps it and the extractors read it back): `read_csv` is wrapped so that
    `groupby(col).apply(f).to_list()` calls f on every group WITH its key column, keys ascending -- pandas < 2.2 behaviour.
Each fixture stores the table text, the parser arguments and the records the reference returns.
"""
import importlib.util
import json
import os
import sys
import types
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "idparse")
REF = "/root/reference/pyascore/parsing/id_parsers.py"

import _msfiles  # noqa: E402
from pyascore_b200.parsing._xml import STD_AA_MASS  # noqa: E402


def load_reference_module():
    pt = types.ModuleType("pyteomics")
    for sub in ("mzid", "pepxml", "mass"):
        m = types.ModuleType("pyteomics." + sub)
        setattr(pt, sub, m)
        sys.modules["pyteomics." + sub] = m
    sys.modules["pyteomics"] = pt
    pt.mzid.MzIdentML = object
    pt.pepxml.PepXML = object
    pt.mass.std_aa_mass = {k: v for k, v in STD_AA_MASS.items() if len(k) == 1}
    spec = importlib.util.spec_from_file_location("ref_id_parsers", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    import pandas

    class _Applied:
        def __init__(self, items):
            self.items = items

        def to_list(self):
            return self.items

    class _Grouped:
        def __init__(self, df, col):
            self.df, self.col = df, col

        def apply(self, func):
            return _Applied([func(g) for _, g in self.df.groupby(self.col)])

    class _Frame:
        def __init__(self, df):
            self.df = df

        def groupby(self, col):
            return _Grouped(self.df, col)

    mod.read_csv = lambda path, sep="\t": _Frame(pandas.read_csv(path, sep=sep))
    return mod


def records_to_json(recs):
    out = []
    for r in recs:
        out.append(dict(scan=int(r["scan"]), charge_state=None if r["charge_state"] is None else int(r["charge_state"]),
                        score=None if r["score"] is None else float(r["score"]), peptide=str(r["peptide"]),
                        mod_positions=[int(x) for x in r["mod_positions"]],
                        mod_masses=[float(x) for x in r["mod_masses"]]))
    return out


def main():
    from test_parsing import QUERIES, load_config1
    ref = load_reference_module()
    _, config1_queries, _ = load_config1()
    extra = [
        dict(scan=7, charge=2, hits=[dict(peptide="CSTC", score="-1.5", nterm="42.0106", mods=[(2, "167.00"), (4, "160.03")]),
                                     dict(peptide="MCK", score="0.25", mods=[(1, "147.04")])]),
        dict(scan=3, charge=4, hits=[dict(peptide="ACS", score="1", mods=[(3, repr(STD_AA_MASS["S"] + 79.9663))])]),
    ]
    cases = {
        "queries_nostatic": (QUERIES, dict(static_mods={})),
        "queries_static_c": (QUERIES, dict(static_mods={"C": 57.021464})),
        "queries_static_nc_threshold": (QUERIES + extra, dict(static_mods={"C": 57.021464, "n": 42.010565}, score_threshold=2.)),
        "extra_static_default": (extra, dict()),
        "config1": (config1_queries, dict(static_mods={"C": 57.021464})),
    }
    os.makedirs(OUT, exist_ok=True)
    for name, (queries, kw) in cases.items():
        for fmt, writer in (("percolatorTXT", _msfiles.write_percolator_txt), ("mokapotTXT", _msfiles.write_mokapot_txt)):
            path = os.path.join(OUT, "_tmp.txt")
            writer(path, queries)
            text = open(path).read()
            mods = dict(ref.COMMON_MODS)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                recs = ref.IdentificationParser(path, fmt, ref.MassCorrector(mod_mass_dict=mods), **kw).to_list()
            os.remove(path)
            json.dump(dict(format=fmt, kwargs=kw, table=text, records=records_to_json(recs)),
                      open(os.path.join(OUT, "%s_%s.json" % (name, fmt)), "w"), indent=1)
            print("%-40s %-14s %3d records" % (name, fmt, len(recs)))


if __name__ == "__main__":
    main()
