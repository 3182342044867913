"""GPU: the secondary public classes (PyBinnedSpectra, PyModifiedPeptide / PyFragmentGraph, PyLogMath,
PyBinomialDist) against the C oracle bit for bit, plus the known answers the reference pins in
test/test_modified_peptide_container.py, test/test_spectra_container.py and test/test_util.py
(m/z literals within the reference's own rtol 1e-6; scipy comparisons within its atol 5e-5)."""
import ctypes as C

import numpy as np
import pytest

import _golden

pytestmark = pytest.mark.gpu

PH = 79.966331


def u32(x):
    return np.array(x, dtype=np.uint32)


@pytest.mark.parametrize("n_top,bin_size", [(40, 100.), (3, 25.), (254, 1000.), (12, 2.5)])
def test_binned_spectra_other_depths_vs_oracle(n_top, bin_size):
    """n_top beyond the scoring depth (exact-rank path of K1 for n_top > 31) and bin counts beyond the
    shared-memory tables (general path), sorted input"""
    from oracle.cscorer import OraclePyAscore
    from pyascore_b200 import PyBinnedSpectra
    rng = np.random.default_rng(77)
    masses = np.sort(rng.uniform(300., 1800., 1200))
    intens = rng.lognormal(5., 1., 1200)
    spec = PyBinnedSpectra(bin_size=bin_size, n_top=n_top)
    spec.consume_spectra(masses, intens)
    ob = OraclePyAscore(bin_size, n_top, "STY", PH).binned(masses, intens)
    k = 0
    for b in range(spec.n_bins):
        spec.bin = b
        for r in range(spec.n_peaks):
            spec.rank = r
            assert (ob["bin"][k], ob["rank"][k], ob["mz"][k], ob["intensity"][k]) == (b, r, spec.mz, spec.intensity)
            k += 1
    assert k == ob["mz"].size


# ---------------------------------------------------------------------------------------------
def test_binned_spectra_toy():
    """test/test_spectra_container.py:15-35"""
    from pyascore_b200 import PyBinnedSpectra
    for kw in (dict(bin_size=100., n_top=10), dict(bin_size=150., n_top=10)):
        assert PyBinnedSpectra(**kw).bin_size == kw["bin_size"]
    masses = np.array([100., 300., 325., 350., 375., 400., 425., 450., 475., 500., 550., 1000.])
    intens = np.array([50., 200., 100., 1000., 500., 100., 1200., 200., 300., 400., 500., 50.])
    spec = PyBinnedSpectra(bin_size=200., n_top=6)
    spec.consume_spectra(masses, intens)
    assert (spec.min_mz, spec.max_mz, spec.n_bins) == (100., 1000., 5)
    got = []
    while spec.bin < spec.n_bins:
        if spec.n_peaks > 0:
            got.append((spec.n_peaks, spec.mz))
        spec.next_bin()
        spec.reset_rank()
    assert got == [(1, 100.), (6, 425.), (2, 550.), (1, 1000.)]
    with pytest.raises(ValueError):
        spec.consume_spectra(masses.astype(np.float32), intens)


def test_binned_spectra_random_vs_numpy_and_oracle():
    """test/test_spectra_container.py:37-67 (unsorted input, possibly negative intensities)"""
    from oracle.cscorer import OraclePyAscore
    from pyascore_b200 import PyBinnedSpectra
    n_top, bin_size, n_peaks = 10, 100., 500
    rng = np.random.RandomState(2345)
    masses = rng.uniform(500., 2000., n_peaks)
    intens = 100. * rng.randn(n_peaks) + 300.
    spec = PyBinnedSpectra(bin_size=bin_size, n_top=n_top)
    spec.consume_spectra(masses, intens)
    lo = np.floor(masses.min() / 100.) * 100.
    assert spec.min_mz == lo and spec.n_bins == int(np.ceil((np.ceil(masses.max() / 100.) * 100. - lo) / bin_size))
    for ind in range(spec.n_bins):
        sel = np.logical_and(masses >= lo + ind * bin_size, masses < lo + (ind + 1) * bin_size)
        order = np.argsort(intens[sel])[::-1]
        bm, bi = masses[sel][order], intens[sel][order]
        assert spec.n_peaks == min(n_top, bm.size)
        for rank in range(min(n_top, bm.size)):
            assert spec.mz == bm[rank] and spec.intensity == bi[rank]
            spec.next_rank()
        spec.next_bin()
        spec.reset_rank()
    ob = OraclePyAscore(bin_size, n_top, "STY", PH).binned(masses, intens)
    spec.reset_bin()
    k = 0
    for b in range(spec.n_bins):
        spec.bin = b
        for r in range(spec.n_peaks):
            spec.rank = r
            assert (ob["bin"][k], ob["rank"][k], ob["mz"][k], ob["intensity"][k]) == (b, r, spec.mz, spec.intensity)
            k += 1
    assert k == ob["mz"].size


# ---------------------------------------------------------------------------------------------
def test_signature_stepping():
    """test/test_modified_peptide_container.py:7-66"""
    from pyascore_b200 import PyModifiedPeptide
    pep = PyModifiedPeptide("STY", PH)
    for args in (("ASK", 1), ("PASSEFK", 2), ("ASK", 1, 1, u32([0]), np.array([20.], np.float32))):
        pep.consume_peptide(*args)
        for t in "by":
            g = pep.get_fragment_graph(t, 1)
            g.incr_signature()
            assert g.is_signature_end()
    pep.consume_peptide("ASTK", 1)
    gb, gy = pep.get_fragment_graph("b", 1), pep.get_fragment_graph("y", 1)
    for sb, sy in (([1, 0], [0, 1]), ([0, 1], [1, 0])):
        assert list(gb.get_signature()) == sb and list(gy.get_signature()) == sy
        gb.incr_signature(), gy.incr_signature()
    assert gb.is_signature_end() and gy.is_signature_end()
    pep.consume_peptide("PASSSSSEFK", 2)
    gb, gy = pep.get_fragment_graph("b", 1), pep.get_fragment_graph("y", 1)
    for sb, sy in (([1, 1, 0, 0, 0], [0, 0, 0, 1, 1]), ([0, 1, 1, 0, 0], [0, 0, 1, 1, 0])):
        assert list(gb.get_signature()) == sb and list(gy.get_signature()) == sy
        for _ in range(4):
            gb.incr_signature(), gy.incr_signature()
    assert not (gb.is_signature_end() or gy.is_signature_end())
    while not gb.is_signature_end() or not gy.is_signature_end():
        gb.incr_signature(), gy.incr_signature()
    assert list(gb.get_signature()) == [0] * 5 and list(gy.get_signature()) == [0] * 5
    assert gb.get_signature().dtype == np.uint64


def test_fragment_known_answers():
    """test/test_modified_peptide_container.py:68-133 (literals, the reference's rtol 1e-6)"""
    from pyascore_b200 import PyModifiedPeptide
    pep = PyModifiedPeptide("STY", PH)
    pep.consume_peptide("PASSSSSEFK", 2)
    truth = {"b": [98.06058, 169.09769, 256.12972, 423.12808, 590.12644, 677.15847, 764.19050, 893.23309, 1040.30150],
             "c": [115.08713, 186.12424, 273.15627, 440.15463, 607.15299, 694.18502, 781.21705, 910.25964, 1057.32805],
             "y": [147.11334, 294.18176, 423.22435, 510.25638, 597.28841, 764.28677, 931.28513, 1018.3171, 1089.3542],
             "z": [130.08680, 277.15521, 406.19780, 493.22983, 580.26186, 747.26022, 914.25858, 1001.29061, 1072.32772]}
    for t, masses in truth.items():
        g = pep.get_fragment_graph(t, 1)
        g.set_signature(u32([0, 1, 1, 0, 0]))
        for m in masses:
            assert np.isclose(g.get_fragment_mz(), m, rtol=1e-6, atol=0)
            g.incr_fragment()
    pep.consume_peptide("ASMTK", 1)
    for mode, frag_lists in (("all", [[71.03711, 238.03547, 369.07596, 470.12364], [71.03711, 158.06914, 289.10963, 470.12364]]),
                             ("reduced", [[71.03711, 238.03547, 369.07596, 470.12364], [158.06914, 289.10963, 470.12364]])):
        g = pep.get_fragment_graph("b", 1, mode=mode)
        seen = 0
        for graph, sig, frags in zip(g.iter_permutations(), [[1, 0], [0, 1]], frag_lists):
            assert list(graph.get_signature()) == sig
            got = list(graph.iter_fragments())
            assert len(got) == len(frags)
            for (mz, label), m in zip(got, frags):
                assert np.isclose(mz, m + 1.007825, rtol=1e-6, atol=0) and label[0] == "b"
            seen += 1
        assert seen == 2
    g = pep.get_fragment_graph("y", 2)
    assert (g.fragment_type, g.charge_state) == ("y", 2)
    assert [g.get_fragment_size(), g.get_fragment_seq()] == [1, "K"]
    g.incr_fragment()
    assert [g.get_fragment_size(), g.get_fragment_seq()] == [2, "KT"]


CASES = [  # scorer kwargs, neutral losses, peptide, k, aux
    (dict(mod_group="STY", mod_mass=PH, mz_error=0.5, fragment_types="by"), [], "MTTTSAAAYGTHLSPHVPHRVLSTSSTLTR", 3, None),
    (dict(mod_group="STY", mod_mass=PH, mz_error=0.02, fragment_types="by"), [("st", 97.9769)], "RPAEATSSPTSPERPR", 2, None),
    (dict(mod_group="STY", mod_mass=PH, mz_error=0.05, fragment_types="cZ"), [("ST", 18.01528), ("st", 97.9769), ("K", 17.026549)],
     "KGPGQPSSPQRK", 1, None),
    (dict(mod_group="K", mod_mass=42.0106, mz_error=0.02, fragment_types="by"), [], "ACKDKCEKR", 2,
     (u32([2, 6]), np.array([57.021464, 57.021464], np.float32))),
    (dict(mod_group="nKc", mod_mass=42.010565, mz_error=0.5, fragment_types="by"), [], "AKSTKR", 2,
     (u32([0]), np.array([10.5], np.float32))),
]


@pytest.mark.parametrize("case", range(len(CASES)))
def test_fragment_graph_and_sdi_vs_oracle(case):
    """every isoform, every fragment (all ion types, charges 0-3) and the site-determining ions of
    isoform pairs: bit-identical to the oracle's FragmentGraph / getSiteDeterminingIons"""
    from oracle.cscorer import OraclePyAscore
    from pyascore_b200 import PyModifiedPeptide
    kw, nls, peptide, k, aux = CASES[case]
    O = OraclePyAscore(100., 10, **kw)
    P = PyModifiedPeptide(**kw)
    for g, m in nls:
        O.add_neutral_loss(g, m)
        P.add_neutral_loss(g, m)
    a = aux if aux else (None, None)
    O.consume_peptide(peptide, k, 3, *a)
    P.consume_peptide(peptide, k, 3, *a)
    S = len(P._sites)
    for t in "bcyzZ":
        for z in (0, 1, 2, 3):
            sig, off, fr = O.fragment_graph(t, z, S)
            g = P.get_fragment_graph(t, z)
            n = 0
            for graph in g.iter_permutations():
                assert list(graph.get_signature()) == list(sig[n]), (t, z, n)
                got = np.array([mz for mz, _ in graph.iter_fragments()], np.float32)
                assert _golden.same_bits(got, fr[off[n]:off[n + 1]]), (t, z, n, got, fr[off[n]:off[n + 1]])
                n += 1
            assert n == sig.shape[0]
    sig, _, _ = O.fragment_graph("b", 1, S)
    rng = np.random.default_rng(case)
    pairs = [(0, 1), (0, sig.shape[0] - 1)] + [tuple(rng.integers(0, sig.shape[0], 2)) for _ in range(6)]
    for i, j in pairs:
        for t in kw["fragment_types"]:
            for zmax in (1, 2, 3):
                ra, rb = O.site_determining(sig[i], sig[j], t, zmax)
                ga, gb = P.get_site_determining_ions(u32(sig[i]), u32(sig[j]), t, zmax)
                assert _golden.same_bits(ga, ra) and _golden.same_bits(gb, rb), (i, j, t, zmax)
    for i in range(min(sig.shape[0], 8)):
        assert P.get_peptide(u32(sig[i])) == O.get_peptide(sig[i])
    assert P.get_peptide() == O.get_peptide(np.zeros(0, np.int32))


def test_log_math_vs_oracle_and_scipy():
    """test/test_util.py:11-72: log_sum (atol 1e-6), binomial functions vs scipy (atol 5e-5, n < 50),
    and every value bit-identical to the oracle's float32 arithmetic"""
    from scipy import special, stats
    from oracle.cscorer import lib
    from pyascore_b200 import PyBinomialDist, PyLogMath
    L = lib("orc_")._dll
    L.orc_log_sum.restype = C.c_float; L.orc_log_sum.argtypes = [C.c_float, C.c_float]
    L.orc_log_bin_coef.restype = C.c_float; L.orc_log_bin_coef.argtypes = [C.c_size_t, C.c_size_t]
    L.orc_binom_new.restype = C.c_void_p; L.orc_binom_new.argtypes = [C.c_float]
    for nm in ("log_pmf", "log_pvalue", "log10_pvalue"):
        f = getattr(L, "orc_binom_" + nm); f.restype = C.c_float; f.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t]
    lm = PyLogMath()
    for a, b in ((-1., -2.), (-30.5, -0.25), (0., 0.), (-np.inf, -3.), (-3., -np.inf), (-700., -1.)):
        got = lm.log_sum(a, b)
        assert np.float32(got).tobytes() == np.float32(L.orc_log_sum(a, b)).tobytes()
        assert np.isclose(got, np.logaddexp(np.float32(a), np.float32(b)), atol=1e-6)
    for n in (1, 2, 7, 20, 49, 300):
        for k in sorted({0, 1, n // 3, n // 2, n - 1, n}):
            got = lm.log_bin_coef(k, n)
            assert np.float32(got).tobytes() == np.float32(L.orc_log_bin_coef(k, n)).tobytes()
            if n < 50:
                assert np.isclose(got, np.log(special.comb(n, k)), atol=5e-5)
    for p in (0.5, 0.01, 0.0004, 0.93):
        d, h = PyBinomialDist(p), L.orc_binom_new(p)
        for n in (1, 5, 20, 49, 120):
            for k in sorted({0, 1, n // 2, n}):
                for nm, sp in (("log_pmf", lambda: stats.binom.logpmf(k, n, np.float32(p))),
                               ("log_pvalue", lambda: stats.binom.logsf(k - 1, n, np.float32(p))),
                               ("log10_pvalue", lambda: stats.binom.logsf(k - 1, n, np.float32(p)) / np.log(10.))):
                    got = getattr(d, nm)(k, n)
                    ref = getattr(L, "orc_binom_" + nm)(h, k, n)
                    assert np.float32(got).tobytes() == np.float32(ref).tobytes(), (p, n, k, nm, got, ref)
                    if n < 50 and p >= 0.01 and np.isfinite(sp()):
                        assert np.isclose(got, sp(), atol=5e-5 * max(1., abs(sp()))), (p, n, k, nm, got, sp())
    with pytest.raises(ValueError):
        d.log_pvalue(5, 3)
