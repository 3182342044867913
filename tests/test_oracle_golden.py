"""CPU: the C restatement (oracle/ascore_oracle.c) against the committed golden vectors that
tests/golden/make_golden.py produced from the unmodified, compiled reference."""
import numpy as np
import pytest

import _golden
from oracle.cscorer import OraclePyAscore
from pyascore_b200 import synth


@pytest.mark.parametrize("name", _golden.golden_names())
def test_oracle_matches_golden(name):
    meta, batch, ref = _golden.load(name)
    if name == "synth_stress":
        pytest.skip("15504-isoform PSMs take ~0.5 s each in the naive oracle; covered by test_oracle_stress_one")
    O = OraclePyAscore(**meta["scorer"])
    for g, m in meta["neutral_losses"]:
        O.add_neutral_loss(g, m)
    n = batch["n_mod"].size
    for i in range(n):
        O.score(*synth.psm_view(batch, i))
        a, b = int(ref["iso_off"][i]), int(ref["iso_off"][i + 1])
        sig, cnt, sc, w, tot = O.pep_score_tables()
        assert w.size == b - a
        assert O.best_sequence == ref["best_sequence"][i]
        assert _golden.same_bits(np.float32(O.best_score), np.float32(ref["best_score"][i]))
        bits = np.zeros(w.size, np.uint64)
        for j in range(sig.shape[1]):
            bits |= sig[:, j].astype(np.uint64) << np.uint64(j)
        assert _golden.same_bits(bits, ref["iso_sig"][a:b])          # same isoforms in the same order
        assert _golden.same_bits(cnt, ref["iso_counts"][a:b])
        assert _golden.same_bits(tot, ref["iso_total"][a:b])
        assert _golden.same_bits(sc, ref["iso_scores"][a:b])
        assert _golden.same_bits(w, ref["iso_weighted"][a:b])
        k = int(batch["n_mod"][i])
        mo = int(ref["mod_off"][i])
        assert _golden.same_bits(O.ascores, ref["ascores"][mo:mo + k])
        alts = O.alt_sites
        for j in range(k):
            assert _golden.same_bits(alts[j], _golden.ref_alt(ref, i, j))


def test_oracle_stress_one():
    meta, batch, ref = _golden.load("synth_stress")
    O = OraclePyAscore(**meta["scorer"])
    O.score(*synth.psm_view(batch, 0))
    b = int(ref["iso_off"][1])
    sig, cnt, sc, w, tot = O.pep_score_tables()
    assert w.size == b == 15504
    assert O.best_sequence == ref["best_sequence"][0]
    assert _golden.same_bits(w, ref["iso_weighted"][:b])             # full std::sort order of 15504 isoforms
    assert _golden.same_bits(cnt, ref["iso_counts"][:b])
    assert _golden.same_bits(O.ascores, ref["ascores"][:5])


@pytest.mark.parametrize("name", _golden.big_names())
def test_oracle_matches_big_golden(name):
    """stress PSMs (15504 isoforms) and all-tie inputs up to 15504 isoforms in "by" and "yb" order: results and the
    SHA-256 of the whole pep_scores table in the reference's listing order (tests/golden/make_golden_big.py)"""
    meta, batch, ref = _golden.load_big(name)
    O = OraclePyAscore(**meta["scorer"])
    for g, m in meta["neutral_losses"]:
        O.add_neutral_loss(g, m)
    n = batch["n_mod"].size
    for i in range(n if not name.startswith("stress") else 10):       # ~0.4 s per stress PSM in the naive oracle
        O.score(*synth.psm_view(batch, i))
        sig, cnt, sc, w, tot = O.pep_score_tables()
        assert w.size == int(ref["n_iso"][i])
        assert O.best_sequence == ref["best_sequence"][i]
        assert _golden.same_bits(np.float32(O.best_score), np.float32(ref["best_score"][i]))
        assert _golden.table_digest(_golden.sig_bits(sig), cnt, sc, w, tot) == ref["table_sha256"][i]
        k = int(batch["n_mod"][i])
        mo = int(ref["mod_off"][i])
        assert _golden.same_bits(O.ascores, ref["ascores"][mo:mo + k])
        for j, alt in enumerate(O.alt_sites):
            assert _golden.same_bits(alt, _golden.ref_alt(ref, i, j))
