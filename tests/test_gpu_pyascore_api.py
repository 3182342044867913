"""GPU: the reference's own test_ascore.py (test/test_ascore.py:10-160), run against the drop-in PyAscore class
on the same 31 fixture PSMs (velos 1/2/3-mod and aux-mod match/spectra pairs, carried in the golden file)."""
import math
import re

import numpy as np
import pytest

import _golden

pytestmark = pytest.mark.gpu


def fixture_psms():
    from pyascore_b200 import synth
    meta, batch, ref = _golden.load("fixtures_by_05")
    assert meta["scorer"]["mod_group"] == "STY"
    return meta, [synth.psm_view(batch, i) for i in range(batch["n_mod"].size)], ref


def scored(meta, psm):
    from pyascore_b200 import PyAscore
    a = PyAscore(**meta["scorer"])
    mz, inten, pep, n_mod, z, aux_pos, aux_mass = psm
    a.score(np.ascontiguousarray(mz), np.ascontiguousarray(inten), pep, n_mod, z,
            np.ascontiguousarray(aux_pos, np.uint32), np.ascontiguousarray(aux_mass, np.float32))
    return a


def test_single_spectrum_score():
    """test/test_ascore.py:10-61: best sequence, score and Ascores of every fixture PSM"""
    meta, psms, ref = fixture_psms()
    for i, psm in enumerate(psms):
        a = scored(meta, psm)
        assert a.best_sequence == ref["best_sequence"][i]
        assert np.float32(a.best_score).tobytes() == np.float32(ref["best_score"][i]).tobytes()
        q = int(ref["mod_off"][i])
        assert _golden.same_bits(np.asarray(a.ascores, np.float32), ref["ascores"][q:q + psm[3]])


def test_alternative_site_consistency():
    """test/test_ascore.py:63-97: alternative sites are free of duplicates and never name a modified site"""
    meta, psms, _ = fixture_psms()
    for psm in psms:
        a = scored(meta, psm)
        for alt in a.alt_sites:
            assert alt.shape[0] == np.unique(alt).shape[0]
        site_iter = re.finditer("[A-Z][^A-Z]*", a.best_sequence)
        modified = [ind + 1 for ind, m in enumerate(site_iter) if "[80]" in m.group()]
        alts = np.concatenate(a.alt_sites) if len(a.alt_sites) else np.zeros(0, np.uint32)
        assert np.intersect1d(modified, alts).shape[0] == 0


def test_pepscore_return():
    """test/test_ascore.py:99-132: one entry per positional isoform, sorted by decreasing weighted score"""
    meta, psms, _ = fixture_psms()
    for psm in psms:
        a = scored(meta, psm)
        ps = a.pep_scores
        nslots, nmods = len(ps[0]["signature"]), psm[3]
        assert len(ps) == math.comb(nslots, nmods)
        assert np.all(np.diff([p["weighted_score"] for p in ps]) <= 0)


def test_ambiguity():
    """test/test_ascore.py:134-160: an isoform is not ambiguous against itself; best vs runner-up equals the
    oracle's value bit for bit and, when the runner-up moves a single modification, is one of the Ascores
    (for two fixture PSMs the runner-up moves both mods and the compiled reference itself misses the Ascores)"""
    from oracle.cscorer import OraclePyAscore
    meta, psms, _ = fixture_psms()
    for psm in psms:
        a = scored(meta, psm)
        ps = a.pep_scores
        assert a.calculate_ambiguity(ps[0], ps[0]) == 0.
        if len(ps) > 1:
            manual = a.calculate_ambiguity(ps[0], ps[1])
            O = OraclePyAscore(**meta["scorer"])
            O.score(*psm)
            ops = O.pep_scores
            assert np.float32(manual).tobytes() == np.float32(O.calculate_ambiguity(ops[0], ops[1])).tobytes()
            moved = int(np.sum(np.asarray(ps[0]["signature"]) != np.asarray(ps[1]["signature"])))
            if moved == 2:
                assert np.any(np.isclose(manual, a.ascores))
