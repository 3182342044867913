"""CPU, only where oracle/_ref (the compiled, unmodified reference) is present: the C restatement
against the real thing on fresh seeded inputs, plus unit checks of the two libstdc++ models."""
import ctypes as C

import numpy as np
import pytest

from oracle import cscorer
from pyascore_b200 import synth

pytestmark = pytest.mark.skipif(not cscorer.available("refshim_"), reason="oracle/_ref not built (needs /root/reference)")


def _same(a, b):
    return a.shape == b.shape and a.tobytes() == b.tobytes()


@pytest.mark.parametrize("workload,n,seed", [("lowres_phospho", 250, 101), ("hires_phospho_nl", 150, 102),
                                             ("acetyl_k", 210, 103)])
def test_oracle_equals_reference_on_synthetic(workload, n, seed):
    w = synth.WORKLOADS[workload]
    batch = synth.make_batch(workload, n, seed=seed, chunk_index=5)
    R = cscorer.RefPyAscore(**w["scorer"])
    O = cscorer.OraclePyAscore(**w["scorer"])
    for g, m in w["neutral_losses"]:
        R.add_neutral_loss(g, m)
        O.add_neutral_loss(g, m)
    for i in range(batch["n_mod"].size):
        a = synth.psm_view(batch, i)
        R.score(*a)
        O.score(*a)
        assert all(_same(x, y) for x, y in zip(R.pep_score_tables(), O.pep_score_tables())), i
        assert R.best_sequence == O.best_sequence and R.sequences() == O.sequences()
        assert _same(R.ascores, O.ascores)
        assert all(_same(x, y) for x, y in zip(R.alt_sites, O.alt_sites))


def test_math_helpers_equal_reference():
    Lr, Lo = cscorer.lib("refshim_"), cscorer.lib("orc_")
    rng = np.random.default_rng(0)
    for a, b in rng.uniform(-60, 0, (200, 2)):
        assert Lr.refshim_log_sum(a, b) == Lo.refshim_log_sum(a, b)
    for n in (1, 2, 7, 34, 68, 312):
        for k in range(0, n + 1, max(1, n // 17)):
            assert Lr.refshim_log_bin_coef(k, n) == Lo.refshim_log_bin_coef(k, n)
    for p in (0.01, 0.0004, 0.1):
        br, bo = Lr.refshim_binom_new(p), Lo.refshim_binom_new(p)
        for n in (5, 34, 150):
            for k in range(1, n + 1, 3):
                assert Lr.refshim_binom_log10_pvalue(br, k, n) == Lo.refshim_binom_log10_pvalue(bo, k, n)
                assert Lr.refshim_binom_log_pmf(br, k, n) == Lo.refshim_binom_log_pmf(bo, k, n)
        Lr.refshim_binom_free(br)
        Lo.refshim_binom_free(bo)


def test_power_set_sum_equals_reference():
    Lr, Lo = cscorer.lib("refshim_"), cscorer.lib("orc_")
    cases = [[], [18.01528], [18.01528, 18.01528], [97.9769, 18.01528, 97.9769], [1., 2., 3., 4.], [0.1, 0.2, 0.3]]
    for v in cases:
        arr = np.array(v, np.float32)
        for depth in (0, 1, 2, 3, 7):       # 0: the reference's unsigned `max_depth - 1` lifts the limit
            a, b = np.zeros(64, np.float32), np.zeros(64, np.float32)
            na = Lr.refshim_power_set_sum(arr.ctypes.data if arr.size else None, arr.size, depth, a.ctypes.data, 64)
            nb = Lo.refshim_power_set_sum(arr.ctypes.data if arr.size else None, arr.size, depth, b.ctypes.data, 64)
            assert na == nb and _same(a[:na], b[:nb]), (v, depth)


def test_all_tie_order_equals_reference():
    """no peak matches -> every isoform scores 0 -> pep_scores lists the raw libstdc++ order
    (hash iteration + introsort on all-equal keys); n = 15, 56, 210, 792"""
    mz, it = np.array([5000., 5001.]), np.array([1., 2.])
    for ft in ("by", "yb"):
        R = cscorer.RefPyAscore(100., 10, "STY", 79.966331, 0.5, ft)
        O = cscorer.OraclePyAscore(100., 10, "STY", 79.966331, 0.5, ft)
        for s, k in ((6, 2), (8, 3), (10, 4), (12, 5)):
            pep = "A" + "S" * s + "K"
            R.score(mz, it, pep, k, 1)
            O.score(mz, it, pep, k, 1)
            assert R.sequences() == O.sequences(), (ft, s, k)
            assert R.best_sequence == O.best_sequence


def test_get_peptide_equals_reference():
    """bracketed sequences incl. the empty-signature rule, short signatures and the overflow mod"""
    for pep, k, aux in (("MTTTSAAAYGTHLSPHVPHRVLSTSSTLTR", 3, None), ("ASK", 3, None),
                        ("KSTMC", 1, (np.array([0, 5], np.uint32), np.array([42.010565, 57.021464], np.float32)))):
        R = cscorer.RefPyAscore(100., 10, "STY", 79.966331)
        O = cscorer.OraclePyAscore(100., 10, "STY", 79.966331)
        a = aux if aux else (None, None)
        R.consume_peptide(pep, k, 1, *a)
        O.consume_peptide(pep, k, 1, *a)
        for sig in ([], [0, 1, 0, 1], [1], [0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1]):
            s = np.array(sig, np.int32)
            assert R.get_peptide(s) == O.get_peptide(s), (pep, sig)


@pytest.mark.parametrize("n_top,bin_size", [(10, 100.), (40, 100.), (3, 25.), (254, 1000.), (12, 2.5)])
def test_binning_equals_reference_on_awkward_intensities(n_top, bin_size):
    """BinnedSpectra on the inputs the GPU binning tests lean on the oracle for: intensities that are distinct
    as doubles but collide as floats, values far outside the float range, negative intensities, many bins"""
    rng = np.random.default_rng(31)
    O = cscorer.OraclePyAscore(bin_size, n_top, "STY", 79.966331)
    R = cscorer.RefPyAscore(bin_size, n_top, "STY", 79.966331)
    n = 900
    mz = np.sort(rng.uniform(300., 1800., n))
    cases = [5000. + rng.permutation(n) * 1e-7,
             np.where(rng.random(n) < 0.3, 321.5 + rng.permutation(n) * 1e-9, rng.lognormal(5., 1., n)),
             rng.standard_normal(n) * 10. ** rng.integers(-60, 60, n)]
    for inten in cases:
        a, b = O.binned(mz, inten), R.binned(mz, inten)
        assert all(np.array_equal(a[k], b[k]) for k in a)


def test_reference_behaviour_for_n_top_above_10():
    """What `n_top > 10` means in the reference, as the reason the B200 library pins n_top to 10 for scoring
    (include/pyascore_b200.h, PA_N_TOP): the weight vector has ten entries (cpp/Ascore.cpp:15-19), so PepScores, the
    isoform order, best_sequence and the alternative sites are those of n_top = 10 -- deeper ranks only add count /
    score columns and can move the depth calculateAmbiguity picks (cpp/Ascore.cpp:165-175)."""
    w = synth.WORKLOADS["lowres_phospho"]
    batch = synth.make_batch("lowres_phospho", 120, seed=7, chunk_index=9)
    kw10, kw12 = dict(w["scorer"]), dict(w["scorer"], n_top=12)
    R10, R12 = cscorer.RefPyAscore(**kw10), cscorer.RefPyAscore(**kw12)
    moved = 0
    for i in range(batch["n_mod"].size):
        a = synth.psm_view(batch, i)
        R10.score(*a)
        R12.score(*a)
        s10, c10, f10, w10, t10 = R10.pep_score_tables()
        s12, c12, f12, w12, t12 = R12.pep_score_tables()
        assert c12.shape[1] == 12 and f12.shape[1] == 12
        assert _same(s10, s12) and _same(w10, w12) and _same(t10, t12)            # same isoforms, order and PepScores
        assert _same(c10, np.ascontiguousarray(c12[:, :10])) and _same(f10, np.ascontiguousarray(f12[:, :10]))
        assert R10.best_sequence == R12.best_sequence
        assert all(_same(x, y) for x, y in zip(R10.alt_sites, R12.alt_sites))
        moved += int(not _same(R10.ascores, R12.ascores))
    assert moved < batch["n_mod"].size                                             # only Ascores can differ, and rarely do
