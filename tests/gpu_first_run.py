"""Ad-hoc driver for the first GPU bring-up: prints detailed mismatches instead of asserting."""
import sys, os, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import _golden
from pyascore_b200 import Scorer, format_results, synth

def run(name):
    meta, batch, ref = _golden.load(name)
    s = Scorer(**meta["scorer"])
    for g, m in meta["neutral_losses"]:
        s.add_neutral_loss(g, m)
    t = time.time(); res = s.score_batch(batch, keep_isoforms=True); dt = time.time() - t
    n = batch["n_mod"].size
    nbad = 0
    for i in range(n):
        seq, best, asc, alts = format_results(s, batch, res, i)
        k = int(batch["n_mod"][i]); mo = int(ref["mod_off"][i])
        a, b = int(ref["iso_off"][i]), int(ref["iso_off"][i + 1])
        sig, cnt, sc, w, tot = s.fetch_pep_scores(i)
        probs = []
        if int(res["psm_status"][i]) != 0: probs.append("status %d" % res["psm_status"][i])
        if int(res["n_iso"][i]) != b - a: probs.append("n_iso %d vs %d" % (res["n_iso"][i], b - a))
        elif not _golden.same_bits(sig, ref["iso_sig"][a:b]): probs.append("order")
        else:
            if not _golden.same_bits(cnt, ref["iso_counts"][a:b]): probs.append("counts")
            if not np.array_equal(tot, ref["iso_total"][a:b]): probs.append("total %s vs %s" % (tot[:3], ref["iso_total"][a:a+3]))
            if not _golden.same_bits(sc, ref["iso_scores"][a:b]): probs.append("scores")
            if not _golden.same_bits(w, ref["iso_weighted"][a:b]): probs.append("weighted")
        if seq != ref["best_sequence"][i]: probs.append("seq %s vs %s" % (seq, ref["best_sequence"][i]))
        if not _golden.same_bits(np.float32(best), np.float32(ref["best_score"][i])): probs.append("best %r vs %r" % (best, float(ref["best_score"][i])))
        if not _golden.same_bits(asc, ref["ascores"][mo:mo + k]): probs.append("asc %s vs %s" % (asc, ref["ascores"][mo:mo + k]))
        for j in range(k):
            if not _golden.same_bits(alts[j], _golden.ref_alt(ref, i, j)): probs.append("alt%d %s vs %s" % (j, alts[j], _golden.ref_alt(ref, i, j)))
        if probs:
            nbad += 1
            if nbad <= 4:
                pep = bytes(batch["pep"][batch["pep_off"][i]:batch["pep_off"][i+1]]).decode()
                print("   PSM %d %s k=%d z=%d: %s" % (i, pep, k, batch["max_charge"][i], "; ".join(probs)))
    print("%-28s psms %4d bad %4d  (%.1f ms) counters %s" % (name, n, nbad, dt * 1e3, {k: v for k, v in s.counters().items() if k.startswith("ms_")}))
    s.close()
    return nbad

if __name__ == "__main__":
    names = sys.argv[1:] or _golden.golden_names()
    tot = 0
    for nm in names:
        try:
            tot += run(nm)
        except Exception:
            traceback.print_exc(); tot += 1
    print("TOTAL BAD", tot)
