"""GPU: the range / asynchronous / multi-device entry points of the C ABI and the sharded CLI.

A batch cut into PSM ranges (pa_score_range), scored asynchronously (pa_score_batch_async + pa_wait) or sharded
over several scorers by `MultiScorer` (pa_shard_ranges: cuts on spectrum boundaries) must give, bit for bit, the
result of one pa_score_batch call over the whole batch.  With one GPU the scorers share it (devices [0, 0]); with
more, every visible GPU takes a range.
"""
import numpy as np
import pytest

import _golden
import _msfiles
from pyascore_b200 import synth
from test_gpu_cli import _argv, _compare
from test_parsing import load_config1

pytestmark = pytest.mark.gpu

WORKLOADS = [("acetyl_k", 2999), ("hires_phospho_nl", 2500), ("lowres_phospho", 4000)]


def _scorer(workload, cls=None, **kw):
    from pyascore_b200 import Scorer
    w = synth.WORKLOADS[workload]
    s = (cls or Scorer)(**w["scorer"], **kw)
    for g, m in w["neutral_losses"]:
        s.add_neutral_loss(g, m)
    return s


def _same(a, b):
    return all(np.asarray(a[k]).tobytes() == np.asarray(b[k]).tobytes() for k in a)


def _devices():
    import torch
    n = torch.cuda.device_count()
    return [[0, 0], [0, 0, 0]] + ([list(range(n))] if n > 1 else [])


@pytest.mark.parametrize("workload,n", WORKLOADS)
def test_ranges_equal_whole(workload, n):
    batch = synth.make_batch(workload, n, seed=99, chunk_index=1)
    s = _scorer(workload)
    whole = s.score_batch(batch)
    npsm = batch["n_mod"].size
    out = {k: np.full_like(v, 0x55) for k, v in whole.items()}
    cuts = [0, npsm // 7, npsm // 7, npsm // 2, npsm]              # an empty range in the middle; cuts may split a spectrum's hits
    for a, b in zip(cuts[:-1], cuts[1:]):
        s.score_batch(batch, out=out, psm_range=(a, b))
        assert s.counters()["n_psm"] == b - a
    assert _same(whole, out)
    with pytest.raises(ValueError):
        s.score_batch(batch, psm_range=(5, npsm + 1))
    s.close()


@pytest.mark.parametrize("workload,n", WORKLOADS)
def test_multiscorer_sharded_equals_unsharded(workload, n):
    from pyascore_b200 import MultiScorer, pin_batch
    batch = synth.make_batch(workload, n, seed=123, chunk_index=2)
    s = _scorer(workload)
    whole = s.score_batch(batch)
    s.close()
    for devices in _devices():
        ms = _scorer(workload, MultiScorer, devices=devices)
        got = ms.score_batch(pin_batch(batch))
        assert _same(whole, got), devices
        ranges = ms.last_ranges
        assert ranges[0][0] == 0 and ranges[-1][1] == batch["n_mod"].size and len(ranges) == len(devices)
        for (a0, a1), (b0, b1) in zip(ranges[:-1], ranges[1:]):
            assert a1 == b0
            if 0 < a1 < batch["n_mod"].size:
                assert batch["psm_spec"][a1] != batch["psm_spec"][a1 - 1]      # a spectrum is binned by one GPU only
        assert sum(c["n_psm"] for c in ms.counters()) == batch["n_mod"].size
        ms.close()


def test_async_and_wait():
    import torch
    batch = synth.make_batch("lowres_phospho", 3000, seed=5, chunk_index=0)
    s = _scorer("lowres_phospho")
    whole = s.score_batch(batch)
    out = s.score_batch_async(batch)
    with pytest.raises(ValueError):
        s.score_batch(batch)                  # one call in flight per scorer
    s.wait()
    s.wait()                                  # nothing in flight: returns at once
    assert _same(whole, out)
    # device-resident inputs produced on the caller's stream: the call must order itself after them
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        dev = {k: torch.from_numpy(v.view(np.int32) if v.dtype == np.uint32 else v).cuda(non_blocking=True) for k, v in batch.items()}
    out_d = s.score_batch_async(dev, stream=st)
    s.wait()
    assert all(out_d[k].cpu().numpy().tobytes() == np.asarray(whole[k]).tobytes() for k in whole)
    s.close()


def test_inconsistent_batches_are_refused():
    """CSR arrays that do not describe a batch come back as PA_ERR_ARG (ValueError), on both input sides"""
    import torch
    base = synth.make_batch("acetyl_k", 300, seed=1)
    s = _scorer("acetyl_k")
    s.score_batch(dict(base))

    def broken(which):
        b = {k: v.copy() for k, v in base.items()}
        if which == "spec_off":
            b["spec_off"][5] = b["spec_off"][4] - 3
        elif which == "pep_off":
            b["pep_off"][10] = b["pep_off"][11] + 4
        elif which == "aux_off":
            b["aux_off"][-1] = -2
        elif which == "mod_off":
            mo = np.zeros(b["n_mod"].size + 1, np.int64)
            np.cumsum(b["n_mod"], out=mo[1:])
            mo[7:] += 1
            b["mod_off"] = mo
        elif which == "aux_null":
            b["aux_pos"] = None
        return b
    for which in ("spec_off", "pep_off", "aux_off", "mod_off", "aux_null"):
        with pytest.raises(ValueError):
            s.score_batch(broken(which))
        if which != "aux_null":
            b = broken(which)
            dev = {k: torch.from_numpy(v.view(np.int32) if v.dtype == np.uint32 else v).cuda() for k, v in b.items()}
            with pytest.raises(ValueError):
                s.score_batch(dev)
    res = s.score_batch(dict(base))           # the scorer is still usable
    assert np.all(res["psm_status"] == 0)
    s.close()


def test_long_peptide_with_neutral_losses_is_scored():
    """fragment budget: counted from the peptide's own neutral-loss stack, not the scorer's worst case
    (a long, high-charge peptide under 3 loss masses used to come back PA_PSM_TOO_MANY_FRAGMENTS)"""
    from oracle.cscorer import OraclePyAscore
    from pyascore_b200 import PyAscore
    rng = np.random.default_rng(3)
    pep = "AGLPEVAGLPEVAGLPEVAGLPEVAGLPEVAGLPEVAGLPEVAGLPEVAGLPEVAGLPEVAGLPSTK"      # 66 residues, losses only at the end
    mz = np.sort(rng.uniform(150., 2000., 400)); inten = rng.lognormal(5., 1., 400)
    kw = dict(bin_size=100., n_top=10, mod_group="STY", mod_mass=79.966331, mz_error=0.05, fragment_types="by")
    a, O = PyAscore(**kw), OraclePyAscore(**kw)
    for sc in (a, O):
        for g, m in (("ST", 18.01528), ("st", 97.9769), ("K", 17.026549)):
            sc.add_neutral_loss(g, m)
    a.score(mz, inten, pep, 1, 4)
    O.score(mz, inten, pep, 1, 4)
    assert a.best_sequence == O.best_sequence
    assert _golden.same_bits(np.float32(a.best_score), np.float32(O.best_score))
    assert _golden.same_bits(a.ascores, O.ascores)


@pytest.mark.parametrize("fmt", ["percolatorTXT", "mokapotTXT"])
def test_config1_cli_bracket_tables(tmp_path, fmt):
    """BASELINE config 1 with the identifications as a percolator / mokapot table: same TSV as the pepXML path
    (the extractors themselves are pinned to the reference's code by tests/golden/idparse)"""
    from pyascore_b200.__main__ import main
    spectra, queries, expected = load_config1()
    spec, ident, out = str(tmp_path / "s.mzML"), str(tmp_path / "i.txt"), str(tmp_path / "o.tsv")
    _msfiles.write_mzml(spec, spectra, inten_bits=32)
    (_msfiles.write_percolator_txt if fmt == "percolatorTXT" else _msfiles.write_mokapot_txt)(ident, queries)
    for setting in ("default", "hit_depth2"):
        main(_argv(expected[setting]["args"]) + ["--ident_file_type", fmt, spec, ident, out])
        _compare(open(out).read(), expected[setting]["tsv"])


def test_config1_cli_sharded(tmp_path):
    """--devices: every chunk sharded over the listed GPUs, byte-identical TSV"""
    import torch
    from pyascore_b200.__main__ import main
    spectra, queries, expected = load_config1()
    spec, ident, out = str(tmp_path / "s.mzML"), str(tmp_path / "i.pep.xml"), str(tmp_path / "o.tsv")
    _msfiles.write_mzml(spec, spectra, inten_bits=32)
    _msfiles.write_pepxml(ident, queries)
    devs = ",".join(str(d) for d in (range(torch.cuda.device_count()) if torch.cuda.device_count() > 1 else (0, 0)))
    for setting in ("default", "hires_nl"):
        main(_argv(expected[setting]["args"]) + ["--devices", devs, spec, ident, out])
        _compare(open(out).read(), expected[setting]["tsv"])


@pytest.mark.parametrize("workload,n", WORKLOADS)
def test_float32_intensities_equal_float64(workload, n):
    """pa_batch.inten32: float32 intensities give, bit for bit, what the same values give as float64 (the ranking
    only compares them); covered on sorted spectra, on a permuted one (general binning path) and with exact ties"""
    batch = synth.make_batch(workload, n, seed=77, chunk_index=4)
    i32 = batch["inten"].astype(np.float32)
    a, b = int(batch["spec_off"][3]), int(batch["spec_off"][4])
    i32[a:a + 6] = i32[a]                                            # ties inside a bin: earlier peak wins, both ways
    perm = np.random.default_rng(1).permutation(b - a)               # spectrum 3 unsorted -> the general path
    batch["mz"][a:b] = batch["mz"][a:b][perm]
    i32[a:b] = i32[a:b][perm]
    b64 = dict(batch, inten=i32.astype(np.float64))
    b32 = {k: v for k, v in batch.items() if k != "inten"}
    b32["inten32"] = i32
    s = _scorer(workload)
    r64 = s.score_batch(b64)
    r32 = s.score_batch(b32)
    assert _same(r64, r32)
    assert s.counters()["bytes_h2d"] < 0.8 * (batch["mz"].nbytes * 2)
    import torch
    dev = {k: torch.from_numpy(v.view(np.int32) if v.dtype == np.uint32 else v).cuda() for k, v in b32.items()}
    rd = s.score_batch(dev)
    assert all(rd[k].cpu().numpy().tobytes() == np.asarray(r64[k]).tobytes() for k in r64)
    s.close()


def test_spectra_block_narrows_float32_intensities(tmp_path):
    """SpectraParser.to_csr hands float32 intensities over when the file's values are float32 (and only then)"""
    from pyascore_b200.parsing import SpectraParser
    spectra, _, _ = load_config1()
    p32, p64 = str(tmp_path / "a.mzML"), str(tmp_path / "b.mzML")
    _msfiles.write_mzml(p32, spectra, inten_bits=32)
    wide = [dict(sp, inten=sp["inten"].astype(np.float64) * (1. + 1e-12)) for sp in spectra]
    _msfiles.write_mzml(p64, wide, inten_bits=64)
    c32 = SpectraParser(p32, "mzML").to_csr()
    c64 = SpectraParser(p64, "mzML").to_csr()
    assert c32.inten is None and c32.inten32.dtype == np.float32
    assert c64.inten32 is None and c64.inten.dtype == np.float64
    assert np.array_equal(c32.inten32.astype(np.float64), SpectraParser(p32, "mzML").to_csr(narrow_intensity=False).inten)


@pytest.mark.parametrize("workload,n", WORKLOADS)
def test_narrowed_mz_equals_exact(workload, n, monkeypatch):
    """host batches whose m/z the library narrows to float32 on the host (pa_narrow_mz) give, bit for bit, the results
    of the float64 path -- including spectra with peaks on and next to bin boundaries (they keep an exact copy), an
    unsorted spectrum, and float32 intensities on top"""
    from pyascore_b200 import Scorer
    batch = synth.make_batch(workload, n, seed=55, chunk_index=6)
    off = batch["spec_off"]
    mz = batch["mz"]
    for q, val in ((2, 700.), (5, np.nextafter(900., 0.)), (7, np.nextafter(500., 1000.))):
        a, b = int(off[q]), int(off[q + 1])
        seg = mz[a:b]
        seg[np.argmin(np.abs(seg - val))] = val                       # a peak on / just below / just above a bin boundary
        mz[a:b] = np.sort(seg)
    a, b = int(off[9]), int(off[10])
    perm = np.random.default_rng(0).permutation(b - a)
    mz[a:b] = mz[a:b][perm]
    batch["inten"][a:b] = batch["inten"][a:b][perm]
    w = synth.WORKLOADS[workload]
    res = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("PA_NARROW", mode)
        s = Scorer(**w["scorer"])
        for g, m in w["neutral_losses"]:
            s.add_neutral_loss(g, m)
        res[mode] = s.score_batch(dict(batch))
        c = s.counters()
        res[mode + "_bytes"], res[mode + "_exact"] = c["bytes_h2d"], c["n_spec_exact"]
        if mode == "1":
            b32 = {k: v for k, v in batch.items() if k != "inten"}
            b32["inten32"] = batch["inten"].astype(np.float32)
            r32 = s.score_batch(b32)
            monkeypatch.setenv("PA_NARROW", "0")
            s0 = Scorer(**w["scorer"])
            for g, m in w["neutral_losses"]:
                s0.add_neutral_loss(g, m)
            assert _same(s0.score_batch(b32), r32)
            s0.close()
        s.close()
    assert _same(res["0"], res["1"])
    assert res["0_exact"] == 0 and 1 <= res["1_exact"] < 0.05 * (off.size - 1)
    assert res["1_bytes"] < 0.8 * res["0_bytes"]


@pytest.mark.parametrize("f32", [False, True])
@pytest.mark.parametrize("workload,n", WORKLOADS)
def test_row_form_binning_equals_topn_kernel(workload, n, f32, monkeypatch):
    """K1 has two kernels: k_bin_rows takes the spectra that are the rule and lists the rest for k_bin_topn.  With the
    row form switched off (PA_K1=topn) k_bin_topn does everything; both ways must agree bit for bit -- on spectra with
    tied intensities inside a bin (exact re-ranking), an unsorted spectrum, NaN / negative / huge m/z, an empty spectrum's
    neighbours and peaks on bin boundaries (spectra beyond the row form's slot: the 2000-peak stress goldens)"""
    batch = synth.make_batch(workload, n, seed=77, chunk_index=3)
    off, mz, inten = batch["spec_off"], batch["mz"], batch["inten"]
    rng = np.random.default_rng(5)
    seg = lambda q: slice(int(off[q]), int(off[q + 1]))
    for q in (1, 4, 11):                                               # ties: whole runs of equal intensities
        sl = seg(q)
        inten[sl] = np.round(inten[sl], -1) + 10.
    v = inten[seg(2)]
    v[1::2] = v[0]                                                     # half the peaks share one intensity
    sl = seg(6)
    perm = rng.permutation(sl.stop - sl.start)
    mz[sl], inten[sl] = mz[sl][perm], inten[sl][perm]                  # unsorted
    mz[int(off[8]) + 3] = np.nan
    mz[int(off[9])] = -5.
    mz[int(off[14]) - 1] = 3.0e6                                       # ascending, but an absurd range
    sl = seg(15)
    mz[sl] = np.sort(np.where(rng.random(sl.stop - sl.start) < 0.2, np.round(mz[sl], -2), mz[sl]))   # peaks on bin boundaries
    v = mz[seg(16)]
    v[1::2] = v[::2][: v[1::2].size]                                   # pairs of equal m/z
    if f32:
        batch = {k: v for k, v in batch.items() if k != "inten"}
        batch["inten32"] = inten.astype(np.float32)
    res = {}
    for mode in ("topn", "rows"):
        monkeypatch.setenv("PA_K1", mode)
        s = _scorer(workload)
        res[mode] = s.score_batch(dict(batch))
        res[mode + "_launches"] = s.counters()["launches_bin"]
        s.close()
    assert _same(res["topn"], res["rows"])
    assert res["rows_launches"] == 2 * res["topn_launches"]


def test_row_form_binning_large_spectra(monkeypatch):
    """spectra of 500 .. 900 peaks (the row form's slot holds up to 1024) and one spectrum whose peaks all fall into one
    bin (more than 511 peaks in a run: declined, the count decode of the row form stops at 511)"""
    from pyascore_b200 import Scorer
    w = dict(synth.WORKLOADS["lowres_phospho"], noise=(500, 900))
    batch = synth.make_batch(w, 1500, seed=5)
    off, mz = batch["spec_off"], batch["mz"]
    a, b = int(off[3]), int(off[4])
    mz[a:b] = np.sort(np.random.default_rng(1).uniform(400.5, 499.5, b - a))
    assert b - a > 511
    res = {}
    for mode in ("topn", "rows"):
        monkeypatch.setenv("PA_K1", mode)
        s = Scorer(**w["scorer"])
        res[mode] = s.score_batch(dict(batch))
        s.close()
    assert _same(res["topn"], res["rows"])


def test_inconsistent_offsets_with_narrowing_forced(monkeypatch):
    """the host narrowing pass works a chunk ahead of that chunk's consistency check: a later chunk whose spectrum offsets
    jump out of the batch and back must be refused (ValueError), not read or written out of bounds by the pass"""
    monkeypatch.setenv("PA_NARROW", "1")
    base = synth.make_batch("lowres_phospho", 60000, seed=4)
    s = _scorer("lowres_phospho")
    good = s.score_batch(dict(base))
    assert s.counters()["n_chunks"] >= 3 and s.counters()["n_spec_exact"] >= 0
    for where in (0.55, 0.9):
        b = {k: v.copy() for k, v in base.items()}
        q = int(where * (b["spec_off"].size - 1))
        b["spec_off"][q] += 10 ** 9
        with pytest.raises(ValueError):
            s.score_batch(b)
    again = s.score_batch(dict(base))           # the scorer is still usable and still right
    assert _same(good, again)
    s.close()
