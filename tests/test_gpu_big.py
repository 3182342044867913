"""GPU: the CUDA library against the goldens whose per-isoform tables are too big to commit
(tests/golden/big, made by tests/golden/make_golden_big.py from the compiled reference):

  * 64 PSMs of the combinatorial stress config (15 504 isoforms each),
  * all-tie and partial-tie inputs with 40 ... 15 504 isoforms in "by" and "yb" order -- the shapes that reach
    the std::sort replay of k_select (shared-memory and global arenas, the rest-list route of k_select_thread),

each checked on best_sequence / best_score / ascores / alt_sites and on the SHA-256 of the whole pep_scores table
in the reference's listing order; plus 16 further stress PSMs against the live C oracle.
"""
import numpy as np
import pytest

import _golden
from pyascore_b200 import synth
from test_gpu_parity import make_scorer, oracle_reference

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", _golden.big_names())
def test_big_golden(name):
    from pyascore_b200 import format_results
    meta, batch, ref = _golden.load_big(name)
    s = make_scorer(meta)
    res = s.score_batch(batch, keep_isoforms=True)
    n = batch["n_mod"].size
    assert np.all(res["psm_status"] == 0), res["psm_status"]
    assert np.array_equal(res["n_iso"], ref["n_iso"])
    bad = []
    for i in range(n):
        seq, best, asc, alts = format_results(s, batch, res, i)
        k = int(batch["n_mod"][i])
        mo = int(ref["mod_off"][i])
        if seq != ref["best_sequence"][i]:
            bad.append((i, "best_sequence", seq, ref["best_sequence"][i]))
        if not _golden.same_bits(np.float32(best), np.float32(ref["best_score"][i])):
            bad.append((i, "best_score", best, float(ref["best_score"][i])))
        if not _golden.same_bits(asc, ref["ascores"][mo:mo + k]):
            bad.append((i, "ascores", asc, ref["ascores"][mo:mo + k]))
        for j in range(k):
            if not _golden.same_bits(alts[j], _golden.ref_alt(ref, i, j)):
                bad.append((i, "alt_sites", j, alts[j], _golden.ref_alt(ref, i, j)))
        sig, cnt, sc, w, tot = s.fetch_pep_scores(i)
        if _golden.table_digest(sig, cnt, sc, w, tot) != ref["table_sha256"][i]:
            bad.append((i, "pep_scores table (order, counts or scores)", int(w.size)))
    assert not bad, "%d mismatches, first: %r" % (len(bad), bad[:5])
    s.close()


def test_big_golden_in_large_batch():
    """the same all-tie PSMs repeated inside one batch of a few hundred PSMs (several per warp of the selection
    kernels, rest list longer than one wave) must give the same best isoform as scored alone"""
    meta, batch, ref = _golden.load_big("ties_by")
    from tests_helpers import repeat_batch
    big, reps = repeat_batch(batch, 24)
    s = make_scorer(meta)
    res = s.score_batch(big)
    one = s.score_batch(batch)
    n = batch["n_mod"].size
    for r in range(reps):
        assert np.array_equal(res["best_sig"][r * n:(r + 1) * n], one["best_sig"])
        assert res["best_score"][r * n:(r + 1) * n].tobytes() == one["best_score"].tobytes()
    assert one["ascores"].tobytes() * reps == res["ascores"].tobytes()
    s.close()


def test_stress_live_oracle():
    """16 stress PSMs of another seed: every result against the C oracle run here"""
    from pyascore_b200 import format_results
    w = synth.WORKLOADS["stress"]
    meta = dict(scorer=w["scorer"], neutral_losses=w["neutral_losses"])
    batch = synth.make_batch("stress", 16, seed=31337, chunk_index=2)
    s = make_scorer(meta)
    res = s.score_batch(batch)
    assert np.all(res["psm_status"] == 0)
    ref = oracle_reference(meta, batch, range(16))
    for i in range(16):
        seq, best, asc, alts = format_results(s, batch, res, i)
        assert seq == ref["best_sequence"][i], (i, seq, ref["best_sequence"][i])
        assert int(res["n_iso"][i]) == ref["n_iso"][i] == 15504
        assert _golden.same_bits(np.float32(best), np.float32(ref["best_score"][i]))
        assert _golden.same_bits(asc, ref["ascores"][i]), (i, asc, ref["ascores"][i])
        assert all(_golden.same_bits(x, y) for x, y in zip(alts, ref["alts"][i]))
    s.close()
