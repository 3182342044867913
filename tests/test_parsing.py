"""CPU: the parsing layer -- stdlib readers, MassCorrector, the
PSM packer and the TSV writer.  Modelled on the reference's test/test_spec_parsers.py and
test/test_id_parsers.py; the reference's example files are used when /root/reference is present
(build container), files written by tests/_msfiles.py everywhere."""
import json
import os

import numpy as np
import pytest

import _msfiles
from pyascore_b200.parsing import (IdentificationParser, MassCorrector, PsmPacker, SpectraParser, iter_batches,
                                   process_mods, write_tsv)
from pyascore_b200.parsing._xml import STD_AA_MASS
from pyascore_b200.parsing.id_parsers import COMMON_MODS
from pyascore_b200.parsing.packer import fragment_charge

HERE = os.path.dirname(os.path.abspath(__file__))
EX = "/root/reference/test/example_inputs"
needs_ref = pytest.mark.skipif(not os.path.isdir(EX), reason="reference example files not on this box")

# test/test_spec_parsers.py:5-8
SCAN_NUMBERS = [14760, 18330, 20462, 21996, 26219, 26962, 27845, 31328, 32257, 35669]
PRECURSOR_MZ = [846.306451825194, 871.696163579367, 858.378601074219, 1116.095703125, 858.408142089844,
                1427.79736328125, 827.992004394531, 1078.430162374319, 1023.712707519531, 885.028560474252]


def load_config1():
    z = np.load(os.path.join(HERE, "golden", "cli", "config1.npz"))
    spectra = []
    for i, scan in enumerate(z["scans"]):
        a, b = z["spec_off"][i], z["spec_off"][i + 1]
        spectra.append(dict(scan=int(scan), precursor_mz=float(z["precursor_mz"][i]),
                            precursor_charge=int(z["precursor_charge"][i]), mz=z["mz"][a:b], inten=z["inten"][a:b]))
    queries = json.loads(bytes(z["queries"]).decode())
    for q in queries:
        for h in q["hits"]:
            h["mods"] = [(p, m) for p, m in h["mods"]]
    return spectra, queries, json.loads(bytes(z["expected"]).decode())


def synth_spectra(rng, n=7):
    out = []
    for i in range(n):
        k = int(rng.integers(1, 400))
        out.append(dict(scan=100 + 7 * (n - i), precursor_mz=float(rng.uniform(400, 1200)),
                        precursor_charge=int(rng.integers(2, 5)), mz=np.sort(rng.uniform(100, 2000, k)),
                        inten=rng.lognormal(5, 1, k).astype(np.float32).astype(np.float64)))
    return out


@pytest.mark.parametrize("fmt,compress,bits", [("mzML", False, 32), ("mzML", True, 64), ("mzXML", False, 64), ("mzXML", True, 64)])
def test_spectra_roundtrip(tmp_path, fmt, compress, bits):
    rng = np.random.default_rng(3)
    spectra = synth_spectra(rng)
    spectra.insert(2, dict(scan=5, ms_level=1, precursor_mz=None, precursor_charge=None,
                           mz=np.array([300., 400.]), inten=np.array([1., 2.])))
    path = str(tmp_path / ("s." + fmt))
    if fmt == "mzML":
        _msfiles.write_mzml(path, spectra, compress=compress, inten_bits=bits)
    else:
        _msfiles.write_mzxml(path, spectra, compress=compress, bits=bits)
    got = SpectraParser(path, fmt).to_list()
    want = sorted([s for s in spectra if s.get("ms_level", 2) == 2], key=lambda s: s["scan"])
    assert [g["scan"] for g in got] == [w["scan"] for w in want]          # sorted by scan, MS1 filtered
    for g, w in zip(got, want):
        assert g["ms_level"] == 2 and g["precursor_mz"] == w["precursor_mz"] and g["precursor_charge"] == w["precursor_charge"]
        assert g["mz_values"].dtype == np.float64 and g["intensity_values"].dtype == np.float64
        assert np.array_equal(g["mz_values"], w["mz"]) and np.array_equal(g["intensity_values"], w["inten"])
    assert len(SpectraParser(path, fmt, ms_level=0).to_list()) == len(spectra)     # 0 = every level
    assert [s["scan"] for s in SpectraParser(path, fmt, ms_level=1).to_list()] == [5]
    d = SpectraParser(path, fmt).to_dict()
    assert sorted(d) == [w["scan"] for w in want] and "scan" not in d[want[0]["scan"]]
    csr = SpectraParser(path, fmt).to_csr(pinned=False)
    assert list(csr.scans) == [w["scan"] for w in want] and csr.spec_off[-1] == sum(w["mz"].size for w in want)
    i = csr.index_of(want[3]["scan"])
    assert np.array_equal(csr.spectrum(i)[0], want[3]["mz"]) and csr.index_of(-12345) == -1
    with pytest.raises(ValueError):
        SpectraParser(path, "mgf")


def test_missing_precursor_charge(tmp_path):
    s = [dict(scan=9, precursor_mz=500.25, precursor_charge=0, mz=np.array([200.]), inten=np.array([3.]))]
    for fmt, wr in (("mzML", _msfiles.write_mzml), ("mzXML", _msfiles.write_mzxml)):
        path = str(tmp_path / ("p." + fmt))
        wr(path, s)
        rec = SpectraParser(path, fmt).to_list()[0]
        # the reference's single try block drops BOTH values when the charge is absent (spec_parsers.py:85-98)
        assert rec["precursor_mz"] is None and rec["precursor_charge"] is None


@needs_ref
def test_reference_example_spectra():
    a = SpectraParser(EX + "/spectra/test_spectra.mzML", "mzML").to_list()
    b = SpectraParser(EX + "/spectra/test_spectra.mzXML", "mzXML").to_list()
    for lst in (a, b):
        assert [s["scan"] for s in lst] == SCAN_NUMBERS
        assert [s["precursor_mz"] for s in lst] == PRECURSOR_MZ
        assert all(s["ms_level"] == 2 and s["precursor_charge"] == 3 and s["mz_values"].size > 0 for s in lst)
    for x, y in zip(a, b):
        assert np.array_equal(x["mz_values"], y["mz_values"]) and np.array_equal(x["intensity_values"], y["intensity_values"])
    spectra, _, _ = load_config1()                   # the committed golden holds exactly these arrays
    for x, g in zip(a, spectra):
        assert np.array_equal(x["mz_values"], g["mz"]) and np.array_equal(x["intensity_values"], g["inten"].astype(np.float64))


# ---- MassCorrector: the cases of test/test_id_parsers.py:11-105 --------------------------------
def test_mass_corrector():
    c = MassCorrector()
    ac, ox, ph = 42.010565, 15.9949, 79.966331
    for i in range(6):
        assert c.correct("X", 0, round(ac, i)) == (('n',), (0,), (ac,))
        assert c.correct("M", 1, round(STD_AA_MASS["M"] + ac, i)) == (('n',), (0,), (ac,))
        assert c.correct("M", 1, round(STD_AA_MASS["M"] + ac + ox, i)) == (('n', 'M'), (0, 1), (ac, ox))
        assert c.correct("S", 5, round(STD_AA_MASS["S"] + ph, i)) == (('S',), (5,), (ph,))
    with pytest.warns(UserWarning):
        res, pos, mass = c.correct("M", 5, STD_AA_MASS["M"] + ph)      # phospho on M is unknown: passed through
    assert res == ("M",) and pos == (5,) and abs(mass[0] - ph) < 1e-9
    pos, mass = c.correct_multiple("MRAMSLVSNEGDSEQNEIR", np.array([1, 5]),
                                   np.array([STD_AA_MASS["M"] + ac + ox, STD_AA_MASS["S"] + ph]))
    assert list(pos) == [0, 1, 5] and list(mass) == [ac, ox, ph]
    assert c.correct("n", 0, 42.01)[1] == (0,) and MassCorrector(n_mod_ind=1).correct("n", 0, 42.01)[1] == (1,)


QUERIES = [
    dict(scan=40, charge=3, hits=[dict(peptide="MRAMSLVSNEGDSEQNEIR", score="2.5", nterm=None,
                                       mods=[(1, "189.05"), (5, "167.00")]),            # n-term acetyl fused with M-ox
                                  dict(peptide="KEESEESDDDMGFGLFD", score="1.25", mods=[(4, "167.00"), (11, "147.04")])]),
    dict(scan=12, charge=2, hits=[dict(peptide="LLVKKIVSLVR", score="3.75", mods=[])]),
    dict(scan=25, charge=2, hits=[dict(peptide="ACDSTYK", score="0.5", nterm="42.01", mods=[(2, "160.03"), (6, "243.03")])]),
]
WANT = [  # scan-sorted; (scan, charge, score, peptide, positions, masses)
    (12, 2, 3.75, "LLVKKIVSLVR", [], []),
    (25, 2, 0.5, "ACDSTYK", [0, 2, 6], [42.010565, 57.021464, 79.966331]),
    (40, 3, 2.5, "MRAMSLVSNEGDSEQNEIR", [0, 1, 5], [42.010565, 15.9949, 79.966331]),
    (40, 3, 1.25, "KEESEESDDDMGFGLFD", [4, 11], [79.966331, 15.9949]),
]


@pytest.mark.parametrize("fmt", ["pepXML", "mzIdentML", "percolatorTXT", "mokapotTXT"])
def test_identification_formats(tmp_path, fmt):
    path = str(tmp_path / "ids")
    kw = {}
    if fmt == "pepXML":
        _msfiles.write_pepxml(path, QUERIES); kw = dict(score_string="xcorr_score")
    elif fmt == "mzIdentML":
        _msfiles.write_mzid(path, QUERIES); kw = dict(score_string="SEQUEST:xcorr")
    elif fmt == "percolatorTXT":
        _msfiles.write_percolator_txt(path, QUERIES); kw = dict(static_mods={})
    else:
        _msfiles.write_mokapot_txt(path, QUERIES); kw = dict(static_mods={})
    got = IdentificationParser(path, fmt, **kw).to_list()
    assert len(got) == len(WANT)
    for g, (scan, z, score, pep, pos, mass) in zip(got, WANT):
        assert g["scan"] == scan and g["peptide"] == pep and g["score"] == score
        assert g["charge_state"] == (None if fmt == "mokapotTXT" else z)
        assert list(g["mod_positions"]) == pos and list(g["mod_masses"]) == mass
    assert len(IdentificationParser(path, fmt, score_threshold=2., **kw).to_list()) == 2   # keeps scores below it
    d = IdentificationParser(path, fmt, **kw).to_dict()
    assert sorted(d) == [12, 25, 40] and d[40]["peptide"] == "KEESEESDDDMGFGLFD"
    with pytest.raises(ValueError):
        IdentificationParser(path, "sqt")


def test_static_mods_in_bracket_tables(tmp_path):
    """percolator/mokapot tables do not annotate static mods: they are re-attached (id_parsers.py:536-553)"""
    path = str(tmp_path / "p.txt")
    q = [dict(scan=3, charge=2, hits=[dict(peptide="ACS", score="1", mods=[(3, repr(STD_AA_MASS["S"] + 79.9663))])])]
    _msfiles.write_percolator_txt(path, q)
    g = IdentificationParser(path, "percolatorTXT", static_mods={"C": 57.021464, "n": 42.010565}).to_list()[0]
    assert list(g["mod_positions"]) == [0, 2, 3] and list(g["mod_masses"]) == [42.010565, 57.021464, 79.966331]


@needs_ref
def test_reference_example_identifications():
    """the 20 literals of test/test_id_parsers.py:158-197 through both example files"""
    _, queries, _ = load_config1()
    pep = IdentificationParser(EX + "/psms/test_psms.pep.xml", "pepXML", score_string="xcorr_score").to_list()
    mzid = IdentificationParser(EX + "/psms/test_psms.mzid", "mzIdentML", score_string="SEQUEST:xcorr").to_list()
    assert len(pep) == 20 and len(mzid) == 20
    flat = [(q["scan"], q["charge"], h) for q in sorted(queries, key=lambda q: q["scan"]) for h in q["hits"]]
    for g, (scan, z, h) in zip(pep, flat):
        assert g["scan"] == scan and g["charge_state"] == z == 3 and g["peptide"] == h["peptide"]
        assert g["score"] == float(h["score"]) and list(g["mod_positions"]) == [p for p, _ in h["mods"]]
        assert set(g["mod_masses"]) <= {79.966331, 15.9949, 57.021464}
    assert pep[0]["score"] == 4.48925829 and list(pep[6]["mod_positions"]) == [17, 21] and list(pep[6]["mod_masses"]) == [15.9949, 79.966331]
    assert list(pep[17]["mod_positions"]) == [7] and list(pep[17]["mod_masses"]) == [57.021464]
    for g, w in zip(mzid[::2], pep[::2]):                     # Tide's mzid collapses localisations
        assert (g["scan"], g["charge_state"], g["score"], g["peptide"]) == (w["scan"], w["charge_state"], w["score"], w["peptide"])
        assert list(g["mod_positions"]) == list(w["mod_positions"]) and list(g["mod_masses"]) == list(w["mod_masses"])


# ---- packing -----------------------------------------------------------------------------------
def test_process_mods_and_charge():
    pos, mass, n = process_mods("STY", 79.966331, 1., False, "ACDSTYK", [0, 2, 4, 6], [42.010565, 57.021464, 79.97, 80.5])
    assert n == 2 and list(pos) == [0, 2] and pos.dtype == np.uint32 and mass.dtype == np.float32
    assert np.array_equal(mass, np.array([42.010565, 57.021464], np.float32))
    _, _, n = process_mods("STY", 79.966331, 1e-3, False, "ACDSTYK", [4], [79.97])
    assert n == 0                                                     # tight tolerance: treated as fixed
    pos, _, n = process_mods("STY", 79.966331, 1., True, "ACDSTYK", [3, 1], [79.966331, 57.021464])
    assert n == 1 and list(pos) == [2]                                # zero_based shifts by one
    _, _, n = process_mods("nK", 42.010565, 1., False, "AKK", [0], [42.01])
    assert n == 1                                                     # position 0 is residue 'n'
    assert fragment_charge(3, 2, 5) == 2 and fragment_charge(0, 4, 5) == 3 and fragment_charge(None, None, 5) == 1
    assert fragment_charge(1, None, 5) == 1 and fragment_charge(6, None, 3) == 3


def test_packer_batches():
    from pyascore_b200.parsing.spec_parsers import SpectraCSR
    rng = np.random.default_rng(0)
    sizes = [5, 0, 9, 4]
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    csr = SpectraCSR(np.array([10, 20, 30, 40]), np.full(4, np.nan), np.array([3, 0, 2, 0], np.int32), off,
                     rng.uniform(100, 900, off[-1]), rng.uniform(1, 9, off[-1]))
    ph = 79.966331
    psms = [dict(scan=10, charge_state=0, peptide="ASTK", mod_positions=[2], mod_masses=[ph]),
            dict(scan=10, charge_state=0, peptide="ASTK", mod_positions=[3], mod_masses=[ph]),
            dict(scan=30, charge_state=4, peptide="CSYK", mod_positions=[1, 2], mod_masses=[57.021464, ph]),
            dict(scan=30, charge_state=4, peptide="CSYK", mod_positions=[1], mod_masses=[57.021464]),
            dict(scan=40, charge_state=None, peptide="MSSK", mod_positions=[1], mod_masses=[15.9949])]
    (scans, b), = list(iter_batches(csr, psms, hit_depth=1))
    assert list(scans) == [10, 30] and list(b["psm_spec"]) == [0, 2] and b["mz"] is csr.mz
    assert bytes(b["pep"]) == b"ASTKCSYK" and list(b["pep_off"]) == [0, 4, 8] and list(b["n_mod"]) == [1, 1]
    assert list(b["max_charge"]) == [2, 3] and list(b["aux_off"]) == [0, 0, 1] and list(b["aux_pos"]) == [1]
    (scans, b), = list(iter_batches(csr, psms, hit_depth=-1))
    assert list(scans) == [10, 10, 30]                                # no variable mod -> not scored
    chunks = list(iter_batches(csr, psms, hit_depth=-1, chunk_psms=1))
    assert [list(s) for s, _ in chunks] == [[10, 10], [30]]           # chunks are cut on scan boundaries
    with pytest.raises(KeyError):
        PsmPacker(csr).add(dict(scan=11, charge_state=2, peptide="ASK", mod_positions=[2], mod_masses=[ph]))


def test_write_tsv_matches_pandas(tmp_path):
    pd = pytest.importorskip("pandas")
    rows = [[14760, "KMS[80]DDEK", float(np.float32(139.62306)), "137.04561", "13"],
            [26962, "KEDS[80]DS[80]K", float(np.float32(72.198036)), "inf;inf", ";"],
            [3, "", -1.0, "inf", ""], [4, "AS[80]K", 0.0, "0.0;5.5877275", "17,18;35"]]
    a, b = str(tmp_path / "a.tsv"), str(tmp_path / "b.tsv")
    write_tsv(a, rows)
    pd.DataFrame(rows, columns=["Scan", "LocalizedSequence", "PepScore", "Ascores", "AltSites"]).to_csv(b, sep="\t", index=False)
    assert open(a).read() == open(b).read()
    write_tsv(a, [])
    pd.DataFrame([], columns=["Scan", "LocalizedSequence", "PepScore", "Ascores", "AltSites"]).to_csv(b, sep="\t", index=False)
    assert open(a).read() == open(b).read()


def test_cli_parser_and_parameter_file(tmp_path):
    from pyascore_b200.config import args_from_file, build_parser, validate_args
    p = build_parser()
    a = p.parse_args(["s.mzML", "i.pep.xml", "o.tsv"])
    assert (a.residues, a.mod_mass, a.mz_error, a.fragment_types, a.max_fragment_charge, a.hit_depth) == ("STY", 79.966331, .5, "by", 5, 1)
    assert (a.static_mod_groups, a.static_mod_masses, a.spec_file_type, a.ident_file_type) == ("C", "57.021464", "mzML", "pepXML")
    assert a.zero_based is False and a.mod_correction_tol == 1. and not a.match_save
    pf = tmp_path / "params.txt"
    pf.write_text("residues = K   # acetyl\nmod_mass=42.0106\n\n# comment\nmz_error = 0.02\nbad line\n")
    assert args_from_file(str(pf)) == ["--residues", "K", "--mod_mass", "42.0106", "--mz_error", "0.02"]
    b = p.parse_args(args_from_file(str(pf)) + ["--mz_error", "0.05", "s", "i", "o"])
    assert b.residues == "K" and b.mod_mass == 42.0106 and b.mz_error == 0.05       # command line wins
    assert p.parse_args(["--zero_based", "False", "s", "i", "o"]).zero_based is True   # the reference's type=bool quirk
    for bad in (["--residues", "SB"], ["--fragment_types", "bx"], ["--max_fragment_charge", "0"]):
        with pytest.raises(ValueError):
            validate_args(p.parse_args(bad + ["s", "i", "o"]))


# ---- PyPowerSetSum: host routine of the library (no GPU needed) --------------------------------
def test_power_set_sum():
    """test/test_util.py:75-96 + the oracle's restatement for random stacks"""
    import ctypes as C
    from oracle.cscorer import lib
    from pyascore_b200 import PyPowerSetSum
    pss = PyPowerSetSum()
    assert not pss.has_next() and pss.get_sum() == 0.

    def drain(p):
        out = [p.get_sum()]
        while p.has_next():
            p.next()
            out.append(p.get_sum())
        return out
    pss = PyPowerSetSum(np.array([1., 2., 3.], np.float32), 2)
    assert pss.has_next() and drain(pss) == [0., 1., 2., 3., 4., 5.]
    pss.reset(np.array([4., 5., 6.], np.float32), 2)
    assert drain(pss) == [0., 4., 5., 6., 9., 10., 11.]
    pss.reset()
    assert pss.get_sum() == 0. and pss.has_next()
    with pytest.raises(StopIteration):
        drain(pss)
        pss.next()
    L = lib("orc_")._dll
    L.orc_power_set_sum.restype = C.c_long
    L.orc_power_set_sum.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_long]
    rng = np.random.default_rng(11)
    for n in range(0, 7):
        for depth in (0, 1, 2, 3, 9):
            t = rng.choice(np.array([18.01528, 97.9769, 17.026549, 79.966331], np.float32), n).astype(np.float32)
            ref = np.zeros(256, np.float32)
            m = L.orc_power_set_sum(t.ctypes.data, n, depth, ref.ctypes.data, 256)
            got = np.array(drain(PyPowerSetSum(t, depth)), np.float32)
            assert got.tobytes() == ref[:m].tobytes(), (n, depth, got, ref[:m])


# ---- percolator / mokapot tables pinned to the reference's own extractors ------------------------
def _idparse_fixtures():
    import glob
    return sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "idparse", "*.json")))


@pytest.mark.parametrize("path", _idparse_fixtures(), ids=lambda p: os.path.basename(p)[:-5])
def test_bracket_tables_match_reference_extractors(tmp_path, path):
    """records of the REFERENCE's PercolatorTXTExtractor / MokapotTXTExtractor + IdentificationParser run on these
    tables (tests/golden/make_idparse.py) -- same records, same order, same dtypes from ours"""
    import json
    import warnings
    fx = json.load(open(path))
    f = tmp_path / "t.txt"
    f.write_text(fx["table"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = IdentificationParser(str(f), fx["format"], MassCorrector(mod_mass_dict=dict(COMMON_MODS)), **fx["kwargs"]).to_list()
    assert len(got) == len(fx["records"])
    for g, w in zip(got, fx["records"]):
        assert g["scan"] == w["scan"] and g["peptide"] == w["peptide"]
        assert g["charge_state"] == w["charge_state"]
        assert g["score"] == w["score"]
        assert [int(x) for x in g["mod_positions"]] == w["mod_positions"]
        assert [float(x) for x in g["mod_masses"]] == w["mod_masses"]          # exact: same float arithmetic
