"""Sharding of a PSM batch over the GPUs of one box.

PSMs are independent and spectra are read-only, so the path shards with NO collective: every rank
scores a contiguous PSM range cut on spectrum boundaries (all hits of one spectrum stay together,
so a spectrum is binned by exactly one GPU) and writes its slice of the result arrays.  Ranges
are balanced by estimated work, not by count: isoforms x fragments per isoform + peaks to bin.
"""
import numpy as np

_LOG_BINOM = None


def _log_binom(n, k):
    from math import lgamma
    global _LOG_BINOM
    if _LOG_BINOM is None:
        _LOG_BINOM = np.array([lgamma(i + 1) for i in range(256)])
    n = np.clip(n, 0, 255)
    k = np.clip(k, 0, n)
    return _LOG_BINOM[n] - _LOG_BINOM[k] - _LOG_BINOM[n - k]


def estimate_work(batch, mod_group, n_types=2):
    """per-PSM work estimate: C(sites, k) * fragments + peaks/hits (vectorised, host only)"""
    pep = np.asarray(batch["pep"])
    pep_off = np.asarray(batch["pep_off"]).astype(np.int64)
    letters = np.frombuffer("".join(c for c in mod_group if c.isupper()).encode(), np.uint8)
    is_site = np.isin(pep, letters)
    csum = np.concatenate([[0], np.cumsum(is_site)])
    S = (csum[pep_off[1:]] - csum[pep_off[:-1]]).astype(np.int64)
    if "n" in mod_group or "c" in mod_group:
        S = S + 1
    L = np.diff(pep_off)
    k = np.asarray(batch["n_mod"]).astype(np.int64)
    iso = np.where(k <= S, np.exp(np.minimum(_log_binom(S, np.minimum(k, S)), 40.)), 0.)
    frag = n_types * np.maximum(L - 1, 1) * np.maximum(np.asarray(batch["max_charge"]), 1)
    spec = np.asarray(batch["psm_spec"]).astype(np.int64)
    peaks = np.diff(np.asarray(batch["spec_off"]))[spec]
    hits = np.bincount(spec, minlength=int(spec.max()) + 1 if spec.size else 1)[spec]
    return iso * frag + 4.0 * peaks / np.maximum(hits, 1)


def shard_ranges(batch, world, work=None, mod_group="STY"):
    """-> list of (p0, p1) PSM ranges, one per rank, contiguous, cut where psm_spec changes"""
    n = int(np.asarray(batch["n_mod"]).size)
    if world <= 1 or n == 0:
        return [(0, n)] + [(n, n)] * (world - 1)
    spec = np.asarray(batch["psm_spec"])
    if np.any(np.diff(spec) < 0):
        raise ValueError("psm_spec must be non-decreasing to shard on spectrum boundaries")
    if work is None:
        work = estimate_work(batch, mod_group)
    cum = np.concatenate([[0.], np.cumsum(work)])
    starts = np.flatnonzero(np.concatenate([[True], np.diff(spec) != 0]))       # legal cut points
    cuts = [0]
    for r in range(1, world):
        target = cum[-1] * r / world
        j = int(np.searchsorted(cum[starts], target))
        j = min(max(j, 0), starts.size - 1)
        if j > 0 and abs(cum[starts[j - 1]] - target) <= abs(cum[starts[j]] - target):
            j -= 1
        cuts.append(max(int(starts[j]), cuts[-1]))
    cuts.append(n)
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


def take_shard(batch, p0, p1):
    """self-contained CSR batch of PSMs [p0, p1) and the spectra they reference"""
    if p1 <= p0:
        z = lambda dt: np.zeros(0, dt)
        return dict(spec_off=np.zeros(1, np.int64), mz=z(np.float64), inten=z(np.float64), psm_spec=z(np.int32),
                    pep_off=np.zeros(1, np.int32), pep=z(np.uint8), n_mod=z(np.int32), max_charge=z(np.int32),
                    aux_off=np.zeros(1, np.int32), aux_pos=z(np.uint32), aux_mass=z(np.float32))
    s0, s1 = int(batch["psm_spec"][p0]), int(batch["psm_spec"][p1 - 1]) + 1
    a, b = int(batch["spec_off"][s0]), int(batch["spec_off"][s1])
    po = batch["pep_off"][p0:p1 + 1]
    ao = batch["aux_off"][p0:p1 + 1]
    return dict(spec_off=(batch["spec_off"][s0:s1 + 1] - a).astype(np.int64),
                mz=np.ascontiguousarray(batch["mz"][a:b]), inten=np.ascontiguousarray(batch["inten"][a:b]),
                psm_spec=(batch["psm_spec"][p0:p1] - s0).astype(np.int32),
                pep_off=(po - po[0]).astype(np.int32), pep=np.ascontiguousarray(batch["pep"][po[0]:po[-1]]),
                n_mod=np.ascontiguousarray(batch["n_mod"][p0:p1]),
                max_charge=np.ascontiguousarray(batch["max_charge"][p0:p1]),
                aux_off=(ao - ao[0]).astype(np.int32), aux_pos=np.ascontiguousarray(batch["aux_pos"][ao[0]:ao[-1]]),
                aux_mass=np.ascontiguousarray(batch["aux_mass"][ao[0]:ao[-1]]))


def gather_results(parts, ranges, n_psm, mod_off):
    """concatenate per-rank result dicts (host arrays) back into whole-batch arrays"""
    out = {}
    for key in parts[0]:
        per_mod = key in ("ascores", "alt_sites")
        total = int(mod_off[-1]) if per_mod else n_psm
        arr = np.zeros(total, parts[0][key].dtype)
        for part, (p0, p1) in zip(parts, ranges):
            lo, hi = (int(mod_off[p0]), int(mod_off[p1])) if per_mod else (p0, p1)
            arr[lo:hi] = part[key][:hi - lo]
        out[key] = arr
    return out


def shard_ranges_native(batch, world, mod_group="STY", n_types=2, nl_variants=1, peak_weight=55., share=None):
    """The library's own cutter (pa_shard_ranges_for: strided-sample cost estimate, ~1 ms per million PSMs) -- what
    `MultiScorer` uses; host arithmetic only, so it runs without a GPU."""
    import ctypes as C
    from . import _lib
    from .batch import _IN_KEYS, _ptr, add_mod_off
    L = _lib.load()
    add_mod_off(batch)
    pb = _lib.PaBatch()
    pb.n_spec, pb.n_psm = int(batch["spec_off"].shape[0]) - 1, int(batch["n_mod"].shape[0])
    for k in _IN_KEYS:
        setattr(pb, k, _ptr(batch.get(k)))
    cuts = np.zeros(world + 1, np.int64)
    sh = None if share is None else np.ascontiguousarray(share, np.float64)
    rc = L.pa_shard_ranges_for(mod_group.encode("utf8"), int(n_types), int(nl_variants), C.byref(pb), int(world),
                               float(peak_weight), None if sh is None else sh.ctypes.data, cuts.ctypes.data)
    if rc != 0:
        raise ValueError("cannot shard this batch (host arrays with non-decreasing psm_spec needed)")
    return [(int(cuts[r]), int(cuts[r + 1])) for r in range(world)]
