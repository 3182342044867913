#!/usr/bin/env python
"""Benchmark of the PTM-localisation scoring path (BASELINE.json: "PSMs scored/sec").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--psms M]

One "step" = one pass of the hot path over one batch of synthetic PSMs of the named workload
(default: BASELINE config 2, 1M low-res ion-trap phospho PSMs per GPU; weak scaling: every rank
scores its own M PSMs).  Prints ONE JSON line on rank 0:

  value    whole-job PSMs/s with the batch already resident in HBM (device pointers through the
           C ABI; results stay in HBM), timed with CUDA events on the library's stream
  e2e      same metric through the public API with HOST (pinned) buffers: H2D copies of every
           input array and D2H of every result inside the timed region
  roofline dominant kernel: algorithmic bytes / CUDA-event duration vs the measured HBM peak
  cpu_baseline  the compiled reference (oracle/_ref) or, failing that, the C port (oracle/),
           timed on this box's host cores on a bounded sample (rank 0, N=1 only)

--impl reference runs ONLY the CPU reference on all host cores and prints the same JSON shape.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from pyascore_b200 import synth  # noqa: E402

GEN_CHUNK = 32768


def _gen_chunk(args):
    workload, n, seed, ci = args
    return synth.make_batch(workload, n, seed=seed, chunk_index=ci)


def generate(workload, n_psm, seed, procs):
    """n_psm PSMs, generated GEN_CHUNK at a time from (seed, chunk index) on `procs` forked workers"""
    hits = synth.WORKLOADS[workload]["hits"]
    chunk = max(GEN_CHUNK // hits * hits, hits)
    jobs = []
    left = n_psm
    ci = 0
    while left > 0:
        m = min(chunk, left)
        jobs.append((workload, m, seed, ci))
        left -= m
        ci += 1
    if procs > 1 and len(jobs) > 1:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(min(procs, len(jobs))) as pool:
            parts = pool.map(_gen_chunk, jobs)
    else:
        parts = [_gen_chunk(j) for j in jobs]
    return synth.concat_batches(parts) if len(parts) > 1 else parts[0]


# ---------------------------------------------------------------------------------------------
# CPU reference arm (test infrastructure used as a reported baseline only)
# ---------------------------------------------------------------------------------------------
def _cpu_worker(args):
    kind, scorer_kw, nls, sub = args
    from oracle import cscorer
    cls = cscorer.RefPyAscore if kind == "reference" else cscorer.OraclePyAscore
    sc = cls(**scorer_kw)
    for g, m in nls:
        sc.add_neutral_loss(g, m)
    t0 = time.perf_counter()
    sc.score_batch(sub, max_k=8, want_ascores=True)
    return time.perf_counter() - t0


def slice_batch(batch, p0, p1):
    """PSMs [p0,p1) with their spectra (psm_spec must be non-decreasing) as a self-contained batch"""
    s0, s1 = int(batch["psm_spec"][p0]), int(batch["psm_spec"][p1 - 1]) + 1
    a, b = int(batch["spec_off"][s0]), int(batch["spec_off"][s1])
    po = batch["pep_off"][p0:p1 + 1]
    ao = batch["aux_off"][p0:p1 + 1]
    return dict(spec_off=(batch["spec_off"][s0:s1 + 1] - a).astype(np.int64), mz=batch["mz"][a:b].copy(),
                inten=batch["inten"][a:b].copy(), psm_spec=(batch["psm_spec"][p0:p1] - s0).astype(np.int32),
                pep_off=(po - po[0]).astype(np.int32), pep=batch["pep"][po[0]:po[-1]].copy(),
                n_mod=batch["n_mod"][p0:p1].copy(), max_charge=batch["max_charge"][p0:p1].copy(),
                aux_off=(ao - ao[0]).astype(np.int32), aux_pos=batch["aux_pos"][ao[0]:ao[-1]].copy(),
                aux_mass=batch["aux_mass"][ao[0]:ao[-1]].copy())


def cpu_kind():
    from oracle import cscorer
    if cscorer.available("refshim_"):
        return "reference"
    if not cscorer.available("orc_"):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    return "port"


def cpu_throughput(workload, batch, n_sample, cores):
    """PSMs/s of the CPU reference over the first n_sample PSMs, `cores` forked workers, one
    scorer object each (the reference is single-threaded and stateful: SURVEY.md section 2.4)"""
    import multiprocessing as mp
    w = synth.WORKLOADS[workload]
    kind = cpu_kind()
    hits = w["hits"]
    n_sample = max(min(n_sample, batch["n_mod"].size) // hits * hits, hits)
    per = max(n_sample // cores // hits * hits, hits)
    jobs = []
    p = 0
    while p < n_sample:
        q = min(p + per, n_sample)
        jobs.append((kind, w["scorer"], w["neutral_losses"], slice_batch(batch, p, q)))
        p = q
    t0 = time.perf_counter()
    if cores > 1 and len(jobs) > 1:
        with mp.get_context("fork").Pool(min(cores, len(jobs))) as pool:
            pool.map(_cpu_worker, jobs)
    else:
        for j in jobs:
            _cpu_worker(j)
    dt = time.perf_counter() - t0
    return n_sample / dt, kind, n_sample, dt


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = []
        mx = None
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for nm, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def bind_near_gpu(index):
    """Best effort: run this rank (and the pinned buffers it allocates: first touch) on the CPUs of the NUMA
    node its GPU hangs off, so that the end-to-end leg does not cross the socket interconnect.  Returns a
    short description for the JSON line; does nothing when the topology is not visible."""
    try:
        r = subprocess.run(["nvidia-smi", "-i", str(index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                           capture_output=True, text=True, timeout=20)
        bus = r.stdout.strip().splitlines()[0].strip().lower()
        if bus.count(":") == 2 and len(bus.split(":")[0]) == 8:
            bus = bus[4:]                                   # 00000000:1B:00.0 -> 0000:1b:00.0
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip())
        if node < 0:
            return "numa node not reported"
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if len(cpus) < 2:
            return "numa node %d has no usable cpus" % node
        os.sched_setaffinity(0, cpus)
        return "bound to numa node %d (%d cpus)" % (node, len(cpus))
    except Exception as e:
        return "not bound (%s)" % type(e).__name__


def retained_peaks(batch, bin_size, n_top, n_sample=2000):
    """Peaks K1 retains (<= n_top per bin), counted exactly on the first n_sample spectra and scaled to the batch."""
    off = batch["spec_off"]
    n_spec = off.size - 1
    m = min(n_sample, n_spec)
    kept = 0
    for q in range(m):
        mz = batch["mz"][off[q]:off[q + 1]]
        if mz.size == 0:
            continue
        lo = np.float32(np.floor(mz.min() / 100.) * 100.)
        bins = np.floor((mz - np.float64(lo)) / np.float64(np.float32(bin_size))).astype(np.int64)
        kept += int(np.minimum(np.bincount(bins - bins.min()), n_top).sum())
    return kept * n_spec // max(m, 1)


def algorithmic_bytes(batch, res_mod_total):
    """16*P/h + L + 8*A + 24 in, 16 + 4*k + 4*sum|alt| out, summed over the batch
    (alt term bounded by one 8-byte mask per mod, which is what the ABI returns)"""
    n = batch["n_mod"].size
    peaks = int(batch["spec_off"][-1])
    return 16 * peaks + int(batch["pep_off"][-1]) + 8 * int(batch["aux_off"][-1]) + 24 * n + 16 * n + 12 * res_mod_total


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="lowres_phospho", choices=sorted(synth.WORKLOADS))
    ap.add_argument("--psms", type=int, default=0, help="PSMs per GPU per step (default: the BASELINE size)")
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--seed", type=int, default=20261017)
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU leg (profiling runs)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    default_psms = {"lowres_phospho": 1000000, "hires_phospho_nl": 1000000, "stress": 1024, "acetyl_k": 999999}
    n_psm = args.psms or default_psms[args.workload]
    cores = os.cpu_count() or 1
    w = synth.WORKLOADS[args.workload]
    cfg = {"workload": "%s: %d synthetic PSMs/GPU/step (BASELINE config %s), seed %d" % (
        args.workload, n_psm, {"lowres_phospho": 2, "hires_phospho_nl": 3, "stress": 4, "acetyl_k": 5}[args.workload],
        args.seed), "scorer": w["scorer"], "neutral_losses": w["neutral_losses"], "psms_per_gpu": n_psm,
        "l2": "inputs (>= 4 GB per step at the default size) exceed the 126 MB L2; no flush needed"}

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        per_step = args.cpu_sample or {"lowres_phospho": 8192, "hires_phospho_nl": 2048, "stress": 1, "acetyl_k": 6144}[args.workload] * cores
        batch = generate(args.workload, min(per_step, n_psm), args.seed, cores)
        vals = []
        for i in range(args.warmup + args.steps):
            v, kind, ns, dt = cpu_throughput(args.workload, batch, per_step, cores)
            if i >= args.warmup:
                vals.append((v, dt))
        value = float(np.mean([v for v, _ in vals]))
        ms = float(np.mean([dt for _, dt in vals]) * 1e3)
        sample = "first %d PSMs of the workload per step, %d forked workers with one scorer each" % (ns, cores)
        print(json.dumps({
            "impl": "reference", "metric": "PSMs scored/sec", "value": value, "unit": "PSM/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": value, "unit": "PSM/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "PSM/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return

    # ------------------------------------------------------------------ our arm
    numa = bind_near_gpu(local) if world > 1 else "single rank: not bound"
    procs = max(1, min(cores // world, len(os.sched_getaffinity(0))))
    t0 = time.time()
    batch = generate(args.workload, n_psm, args.seed + 1000 * rank, procs)   # before any CUDA call (fork)
    gen_s = time.time() - t0
    n_psm = int(batch["n_mod"].size)

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from pyascore_b200 import Scorer, batch as pb

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    scorer = Scorer(device=local, **w["scorer"])
    for g, m in w["neutral_losses"]:
        scorer.add_neutral_loss(g, m)
    pb.add_mod_off(batch)
    mod_total = int(batch["mod_off"][-1])

    # device-resident copy (value) and pinned host copy (e2e)
    dev = {k: torch.from_numpy(v.view(np.int32) if v.dtype == np.uint32 else v).cuda() for k, v in batch.items()}
    host = pb.pin_batch(batch)
    out_host = {k: pb.pinned_empty(mod_total if k in ("ascores", "alt_sites") else n_psm, dt)
                for k, dt in pb._OUT_DTYPES.items()}
    tm = dict(best_sig=torch.int64, best_score=torch.float32, n_iso=torch.int64, n_sites=torch.int32,
              ascores=torch.float32, alt_sites=torch.int64, psm_status=torch.int32)
    out_dev = {k: torch.empty(max(mod_total if k in ("ascores", "alt_sites") else n_psm, 1), dtype=tm[k], device="cuda")
               for k in tm}

    def run_steps(inputs, outputs, steps):
        ms = 0.0
        ctr = None
        for _ in range(steps):
            scorer.score_batch(inputs, out=outputs)
            ctr = scorer.counters()
            ms += ctr["ms_total"]
        return ms, ctr

    def h2d_peak_gbs(nbytes=1 << 30, reps=4):
        """plain pinned-host -> device copy rate of this box (what bounds the e2e leg)"""
        src = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        dst = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        best = 0.0
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); dst.copy_(src, non_blocking=True); e1.record(); torch.cuda.synchronize()
            best = max(best, nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9)
        return best

    pcie_gbs = h2d_peak_gbs() if not args.no_e2e else None
    sampler = ClockSampler(local)
    # ---- value: inputs resident in HBM ----
    run_steps(dev, out_dev, args.warmup)
    barrier()
    sampler.start()
    t_wall = time.perf_counter()
    ms_dev, ctr_dev = run_steps(dev, out_dev, args.steps)
    barrier()
    wall_dev = time.perf_counter() - t_wall
    # ---- e2e: host buffers through the public API ----
    e2e_steps = 1 if args.no_e2e else args.steps
    run_steps(host, out_host, 1 if args.no_e2e else max(1, min(args.warmup, 2)))
    barrier()
    t_wall = time.perf_counter()
    ms_e2e, ctr_e2e = run_steps(host, out_host, e2e_steps)
    ms_e2e *= args.steps / e2e_steps
    barrier()
    wall_e2e = time.perf_counter() - t_wall
    clocks = sampler.stop()

    # parity guard: both paths must agree bit for bit, and every PSM must have been scored
    same = all(out_dev[k][:out_host[k].size].cpu().numpy().tobytes() == out_host[k].tobytes() for k in out_host)
    n_bad = int((out_host["psm_status"] != 0).sum())

    t = torch.tensor([ms_dev, ms_e2e, wall_dev * 1e3, wall_e2e * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev_max, ms_e2e_max, wall_dev_ms, wall_e2e_ms = [float(x) for x in t.cpu()]
    total_psm = n_psm * world * args.steps

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak = float(json.load(open(peaks_path))["hbm_gbs"])
            peak_src = "measured (MEASURED_PEAKS.json)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        kern = {"bin_topn": ctr_dev["ms_bin"], "plan+scan": ctr_dev["ms_plan"], "count_score": ctr_dev["ms_count"],
                "select": ctr_dev["ms_select"], "ascore": ctr_dev["ms_ascore"]}
        dom = max(("bin_topn", "count_score", "select", "ascore"), key=lambda k: kern[k])
        n_launch = {"bin_topn": ctr_dev["launches_bin"], "count_score": ctr_dev["launches_count"],
                    "select": ctr_dev["launches_select"], "ascore": max(ctr_dev["launches_ascore"] // 2, 1)}[dom]
        peaks_n = int(batch["spec_off"][-1])
        # algorithmic bytes of each kernel's own stage per step (DESIGN.md section "kernels")
        n_spec = batch["spec_off"].size - 1
        hits = w["hits"]
        # K1 writes, and K2 / K3b read, 8 B per retained peak plus the 256-cell index, its header and the count
        retained = 8 * retained_peaks(batch, w["scorer"]["bin_size"], w["scorer"]["n_top"])
        index = 268 * n_spec
        alg = {"bin_topn": 16 * peaks_n + retained + index,
               "count_score": (retained + index) * hits + int(batch["pep_off"][-1]) + 8 * int(batch["aux_off"][-1]) + 12 * n_psm + 24 * int(ctr_dev["n_isoforms"]),
               "select": 4 * int(ctr_dev["n_isoforms"]) + 28 * n_psm + int(batch["pep_off"][-1]) + 26 * mod_total,
               "ascore": (retained + index) * hits + int(batch["pep_off"][-1]) + 8 * int(batch["aux_off"][-1]) + 16 * mod_total + 4 * mod_total}
        # DRAM traffic of the dominant kernel from the committed ncu --set full capture (bytes per PSM at
        # 262144 PSMs/launch, profiles/traffic.json), scaled to this run's PSMs per launch
        traffic = None
        issue = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = tj[args.workload][dom]["dram_bytes_per_psm"] * n_psm / max(n_launch, 1)
            # issue-slot view of the same kernel (the path is bound by warp-instruction issue):
            # executed warp instructions per PSM from the committed ncu capture x PSMs per launch / the launch
            # time measured here, against SMs x 4 schedulers x the SM clock sampled during this run
            props = torch.cuda.get_device_properties(local)
            mhz = clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0
            ipeak = props.multi_processor_count * 4 * mhz * 1e6
            iach = tj[args.workload][dom]["warp_inst_per_psm"] * n_psm / (kern[dom] * 1e-3)
            all_inst = sum(tj[args.workload][k]["warp_inst_per_psm"] for k in ("bin_topn", "count_score", "select", "ascore"))
            issue = {"achieved": iach, "peak": ipeak, "unit": "warp-inst/s", "frac": iach / ipeak,
                     "warp_inst_per_psm": tj[args.workload][dom]["warp_inst_per_psm"],
                     "step_frac": all_inst * n_psm / (sum(kern[k] for k in ("bin_topn", "count_score", "select", "ascore")) * 1e-3) / ipeak,
                     "source": "profiles/traffic.json (ncu smsp__inst_executed.sum) x this run's CUDA-event times"}
        except Exception:
            pass
        dom_ms = kern[dom] / max(n_launch, 1)
        achieved = alg[dom] / max(n_launch, 1) / (dom_ms * 1e-3) / 1e9
        step_alg = algorithmic_bytes(batch, mod_total)
        kernel_ms = sum(kern.values())
        out = {
            "metric": "PSMs scored/sec", "value": total_psm / (ms_dev_max * 1e-3), "unit": "PSM/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg, "clocks": clocks,
            "e2e": {"value": total_psm / (ms_e2e_max * 1e-3), "unit": "PSM/s",
                    "h2d_bytes_per_step": int(ctr_e2e["bytes_h2d"]), "d2h_bytes_per_step": int(ctr_e2e["bytes_d2h"]),
                    "ms_per_step": ms_e2e_max / args.steps, "wall_ms_per_step": wall_e2e_ms / args.steps,
                    "h2d_achieved_gbs": ctr_e2e["bytes_h2d"] / (ms_e2e_max / args.steps * 1e-3) / 1e9,
                    "h2d_copy_peak_gbs": pcie_gbs,
                    "note": "bound by the host->device link: h2d_achieved_gbs vs a plain pinned 1 GiB copy on this box"},
            "gpu_launches": int(ctr_dev["kernel_launches"]) * args.steps,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "note": "path is issue/latency bound, not HBM bound (DESIGN.md); frac is of the HBM copy peak",
                         "algorithmic_bytes_per_launch": alg[dom] / max(n_launch, 1), "launch_ms": dom_ms,
                         "issue": issue,
                         "step": {"algorithmic_bytes": step_alg, "kernel_ms": kernel_ms,
                                  "achieved": step_alg / (kernel_ms * 1e-3) / 1e9,
                                  "frac": step_alg / (kernel_ms * 1e-3) / 1e9 / peak}},
            "kernel_ms_per_step": kern, "wall_ms_per_step": wall_dev_ms / args.steps,
            "isoforms_per_step": int(ctr_dev["n_isoforms"]), "fragment_lookups_per_step": int(ctr_dev["n_fragment_lookups"]),
            "host_and_device_paths_bit_identical": bool(same), "psms_not_scored": n_bad, "gen_seconds": gen_s,
            "numa": numa,
        }
        if world == 1 and not args.no_cpu_baseline:
            # CPU reference on this box's host cores, in a fresh process (no fork after CUDA init)
            cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "0",
                   "--workload", args.workload, "--seed", str(args.seed), "--psms", str(n_psm)]
            if args.cpu_sample:
                cmd += ["--cpu-sample", str(args.cpu_sample)]
            try:
                r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
                out["cpu_baseline"] = json.loads(r.stdout.strip().splitlines()[-1])["cpu_baseline"]
            except Exception as e:  # the baseline is reported, never required
                out["cpu_baseline"] = {"value": None, "unit": "PSM/s", "cores": cores, "kind": "unavailable",
                                       "sample": "failed: %r" % (e,)}
        print(json.dumps(out))
    scorer.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
