"""Golden vectors for the PSMs whose per-isoform tables are too big to commit: the combinatorial stress config
(15 504 isoforms per PSM) and all-tie inputs that exercise the reference's tie order (libstdc++ hash-iteration order
+ introsort) at 210 ... 15 504 isoforms.  Generated from the UNMODIFIED reference (oracle/_ref):

    python tests/golden/make_golden_big.py

Per PSM the file keeps the reference's best_sequence / best_score / ascores / alt_sites / n_iso and a SHA-256 over
its whole pep_scores table IN THE REFERENCE'S LISTING ORDER (signature bits u64 | counts i32 | scores f32 | weighted
f32 | total_fragments i32), so that the isoform order, every count and every score are still checked bit for bit.
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.cscorer import RefPyAscore  # noqa: E402
from pyascore_b200 import synth  # noqa: E402
from make_golden import batch_from_psms  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "big")


def table_digest(bits, cnt, sc, w, tot):
    h = hashlib.sha256()
    for a, dt in ((bits, np.uint64), (cnt, np.int32), (sc, np.float32), (w, np.float32), (tot, np.int32)):
        h.update(np.ascontiguousarray(a, dt).tobytes())
    return h.hexdigest()


def sig_bits(sig):
    bits = np.zeros(sig.shape[0], np.uint64)
    for j in range(sig.shape[1]):
        bits |= sig[:, j].astype(np.uint64) << np.uint64(j)
    return bits


def run_reference(scorer_kw, neutral_losses, batch):
    R = RefPyAscore(**scorer_kw)
    for g, m in neutral_losses:
        R.add_neutral_loss(g, m)
    n = batch["n_mod"].size
    seqs, digests = [], []
    best = np.zeros(n, np.float32); n_iso = np.zeros(n, np.int64)
    asc = []; alt_off = [0]; alt = []
    for i in range(n):
        R.score(*synth.psm_view(batch, i))
        seqs.append(R.best_sequence); best[i] = R.best_score
        sig, cnt, sc, w, tot = R.pep_score_tables()
        n_iso[i] = w.size
        digests.append(table_digest(sig_bits(sig), cnt, sc, w, tot))
        asc.append(R.ascores)
        for s in R.alt_sites:
            alt.append(np.asarray(s, np.uint32)); alt_off.append(alt_off[-1] + len(s))
    return dict(best_sequence=np.frombuffer("\n".join(seqs).encode(), np.uint8), best_score=best, n_iso=n_iso,
                ascores=np.concatenate(asc) if asc else np.zeros(0, np.float32), alt_off=np.array(alt_off, np.int64),
                alt=np.concatenate(alt) if alt else np.zeros(0, np.uint32),
                table_sha256=np.frombuffer("\n".join(digests).encode(), np.uint8))


def save(name, scorer_kw, neutral_losses, batch):
    ref = run_reference(scorer_kw, neutral_losses, batch)
    meta = dict(scorer=scorer_kw, neutral_losses=neutral_losses)
    arrays = {"in_" + k: v for k, v in batch.items()}
    arrays.update({"ref_" + k: v for k, v in ref.items()})
    arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), np.uint8)
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("%-24s psms %4d isoforms %8d  %.0f KB" % (name, batch["n_mod"].size, int(ref["n_iso"].sum()), os.path.getsize(path) / 1e3))


def tie_cases():
    """no fragment can match ([5000, 5001] Th): every isoform scores 0 and the listing order is the raw tie order;
    plus sparse spectra where only a handful of low fragments match, so that the TOP score is tied among many
    isoforms but not all (the sort has real work to do around the ties)"""
    far = (np.array([5000., 5001.]), np.array([1., 2.]))
    rng = np.random.default_rng(11)
    P = []
    for n_s, k in ((10, 4), (12, 5), (18, 4), (20, 5), (11, 5), (12, 4), (10, 3), (9, 4), (24, 2), (30, 2), (40, 1)):
        P.append(dict(mz=far[0], inten=far[1], peptide="A" + "S" * n_s + "K", n_mod=k, max_charge=1))
    # partial ties: a few peaks near the first b / last y ions of "A" + S...: isoforms that agree on the first sites tie
    for n_s, k, npk in ((10, 4, 3), (12, 5, 4), (12, 4, 2), (11, 5, 3), (18, 4, 5), (20, 5, 6), (20, 5, 12), (16, 6, 8)):
        pep = "A" + "S" * n_s + "K"
        mz = np.sort(rng.uniform(150., 87.03 * n_s + 300., npk))
        P.append(dict(mz=mz, inten=np.exp(rng.normal(5., 1., npk)), peptide=pep, n_mod=k, max_charge=1))
    # mixed residues so that the hash order is not the one of a homopolymer
    P.append(dict(mz=far[0], inten=far[1], peptide="STYSTYSTYSTYSTYSTYK", n_mod=4, max_charge=2))
    P.append(dict(mz=far[0], inten=far[1], peptide="KSASTSYSASTSYSASTSYSAR", n_mod=3, max_charge=1))
    return batch_from_psms(P)


def main():
    phos = dict(bin_size=100., n_top=10, mod_group="STY", mod_mass=79.966331, mz_error=0.5, fragment_types="by")
    w = synth.WORKLOADS["stress"]
    save("stress_64", w["scorer"], [], synth.make_batch("stress", 64, seed=777, chunk_index=1))
    ties = tie_cases()
    save("ties_by", phos, [], ties)
    save("ties_yb", dict(phos, fragment_types="yb"), [], ties)


if __name__ == "__main__":
    main()
