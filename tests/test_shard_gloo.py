"""CPU, world_size 2 over gloo: the N>1 path of bench.py / a multi-GPU caller.  Each rank takes its
shard (pyascore_b200.shard), scores it -- here with the C oracle standing in for the GPU scorer --
and rank 0 reassembles the slices; the result must equal scoring the whole batch in one go."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from pyascore_b200 import shard, synth  # noqa: E402


def _score_with_oracle(batch, workload):
    from oracle.cscorer import OraclePyAscore
    w = synth.WORKLOADS[workload]
    O = OraclePyAscore(**w["scorer"])
    n = batch["n_mod"].size
    best = np.zeros(n, np.float32)
    asc = []
    for i in range(n):
        O.score(*synth.psm_view(batch, i))
        best[i] = O.best_score
        asc.append(O.ascores)
    return dict(best_score=best, ascores=np.concatenate(asc) if asc else np.zeros(0, np.float32))


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    batch = synth.make_batch("acetyl_k", 240, seed=11, chunk_index=0)     # 3 hits per spectrum
    ranges = shard.shard_ranges(batch, world, mod_group="K")
    p0, p1 = ranges[rank]
    part = _score_with_oracle(shard.take_shard(batch, p0, p1), "acetyl_k")
    dist.barrier()
    # timing max-over-ranks like bench.py
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gathered = [None] * world
    dist.all_gather_object(gathered, part)
    if rank == 0:
        q.put((ranges, gathered, float(t.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    ranges, parts, tmax = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tmax == 2.0
    batch = synth.make_batch("acetyl_k", 240, seed=11, chunk_index=0)
    mod_off = np.concatenate([[0], np.cumsum(batch["n_mod"])]).astype(np.int64)
    # ranges partition the PSMs and never split a spectrum
    assert ranges[0][0] == 0 and ranges[-1][1] == batch["n_mod"].size
    for (a0, a1), (b0, b1) in zip(ranges[:-1], ranges[1:]):
        assert a1 == b0
        if 0 < a1 < batch["n_mod"].size:
            assert batch["psm_spec"][a1] != batch["psm_spec"][a1 - 1]
    whole = _score_with_oracle(batch, "acetyl_k")
    got = shard.gather_results(parts, ranges, batch["n_mod"].size, mod_off)
    assert got["best_score"].tobytes() == whole["best_score"].tobytes()
    assert got["ascores"].tobytes() == whole["ascores"].tobytes()


def test_shard_balance_by_work():
    batch = synth.make_batch("lowres_phospho", 4000, seed=3, chunk_index=0)
    work = shard.estimate_work(batch, "STY")
    for world in (2, 4, 8):
        ranges = shard.shard_ranges(batch, world, work=work)
        loads = np.array([work[a:b].sum() for a, b in ranges])
        assert loads.min() > 0.8 * loads.mean() and loads.max() < 1.2 * loads.mean()
        assert sum(b - a for a, b in ranges) == 4000
