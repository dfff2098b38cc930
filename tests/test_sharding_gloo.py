"""CPU, world_size 2 over gloo: the multi-GPU path is "independent pairs sharded over ranks, no data-path
collective".  The only distributed logic is the shard assignment and the max-over-ranks timing reduction;
both are exercised here with the CPU oracle standing in for the device (the GPU library is not involved)."""
import os
import socket
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank: int, world: int, port: int, n_frames: int, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    from pyoracle import Oracle, options
    from soc_project_stereo_matching_b200.sharding import shard_range
    from soc_project_stereo_matching_b200.synth import make_pair

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lo, hi = shard_range(n_frames, rank, world)
    orc = Oracle()
    opts = options(max_disparity=16, num_paths=4)
    sums = []
    for k in range(lo, hi):
        l, r, _ = make_pair(40, 24, 16, seed=0xB200 + k, texture="scene")
        d = orc.match(l, r, opts, stages=False)["disp_final"]
        sums.append((k, float(np.where(np.isfinite(d), d, 0).sum())))
    # timing reduction used by bench.py: max over ranks
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gathered = [None] * world
    dist.all_gather_object(gathered, sums)
    if rank == 0:
        q.put((float(t.item()), gathered))
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process():
    from pyoracle import Oracle, options
    from soc_project_stereo_matching_b200.sharding import shard_range
    from soc_project_stereo_matching_b200.synth import make_pair

    n = 5
    # contiguous shards covering every frame exactly once (same rule as SGMB_MatchBatchMultiGPU)
    for world in (1, 2, 3, 4, 8):
        covered = [k for r in range(world) for k in range(*shard_range(n, r, world))]
        assert covered == list(range(n))
    assert [shard_range(256, r, 8) for r in range(8)] == [(32 * r, 32 * r + 32) for r in range(8)]

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    tmax, gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tmax == 2.0
    got = dict(x for part in gathered for x in part)
    orc = Oracle()
    opts = options(max_disparity=16, num_paths=4)
    for k in range(n):
        l, r, _ = make_pair(40, 24, 16, seed=0xB200 + k, texture="scene")
        d = orc.match(l, r, opts, stages=False)["disp_final"]
        assert got[k] == float(np.where(np.isfinite(d), d, 0).sum())
