"""CPU: the arithmetic bench.py reports with (algorithmic bytes, retained peaks, batch slicing) against the oracle."""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_retained_peaks_matches_oracle_binning():
    """roofline bytes use the number of peaks K1 retains: bench.py's host count must equal the oracle's BinnedSpectra"""
    from oracle.cscorer import OraclePyAscore
    from pyascore_b200 import synth
    bench = _bench()
    for workload in ("lowres_phospho", "acetyl_k"):
        w = synth.WORKLOADS[workload]
        batch = synth.make_batch(workload, 96 // w["hits"] * w["hits"], seed=7, chunk_index=1)
        n_spec = batch["spec_off"].size - 1
        O = OraclePyAscore(**w["scorer"])
        want = 0
        for q in range(n_spec):
            a, b = int(batch["spec_off"][q]), int(batch["spec_off"][q + 1])
            want += O.binned(batch["mz"][a:b], batch["inten"][a:b])["mz"].size
        got = bench.retained_peaks(batch, w["scorer"]["bin_size"], w["scorer"]["n_top"], n_sample=n_spec)
        assert got == want, (workload, got, want)


def test_slice_batch_is_self_contained():
    from pyascore_b200 import synth
    bench = _bench()
    batch = synth.make_batch("acetyl_k", 60, seed=3, chunk_index=0)
    sub = bench.slice_batch(batch, 9, 30)
    assert sub["n_mod"].size == 21 and sub["spec_off"][0] == 0 and sub["pep_off"][0] == 0
    assert sub["psm_spec"].min() == 0 and sub["psm_spec"].max() == sub["spec_off"].size - 2
    for i in (0, 7, 20):
        a = synth.psm_view(batch, 9 + i)
        b = synth.psm_view(sub, i)
        for x, y in zip(a, b):
            assert np.array_equal(np.asarray(x), np.asarray(y)) if not isinstance(x, (str, int)) else x == y


def test_algorithmic_bytes_formula():
    """16 P/h + L + 8 A + 24 in, 16 + 12 per mod out"""
    from pyascore_b200 import synth
    bench = _bench()
    batch = synth.make_batch("lowres_phospho", 32, seed=5, chunk_index=0)
    n = batch["n_mod"].size
    mods = int(batch["n_mod"].sum())
    want = 16 * int(batch["spec_off"][-1]) + int(batch["pep_off"][-1]) + 8 * int(batch["aux_off"][-1]) + 40 * n + 12 * mods
    assert bench.algorithmic_bytes(batch, mods) == want
