"""CPU, world_size 2 over gloo: the multi-GPU path is "independent pairs sharded over devices, no data-path collective".
What is distributed is (a) the shard assignment - the C++ rule inside libsgm_b200.so that SGMB_MatchBatchMultiGPU and
SGMB_PoolMatchBatch apply, reached here through its host-only export SGMB_ShardRange (no CUDA call, so it runs without a
GPU) - and (b) bench.py's max-over-ranks timing reduction.  Each rank asks the library for its shard; together the ranks
must cover every pair exactly once, contiguously, and agree with the closed form documented in include/sgm_b200.h."""
import os
import socket
import sys

import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank: int, world: int, port: int, sizes, q):
    sys.path.insert(0, ROOT)
    import torch
    import soc_project_stereo_matching_b200 as sgm

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    mine = [sgm.shard_range(n, world, rank) for n in sizes]            # the library's own rule (C++), one call per batch size
    # timing reduction used by bench.py: max over ranks
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        q.put((float(t.item()), gathered))
    dist.destroy_process_group()


def test_shard_rule_of_the_library():
    import soc_project_stereo_matching_b200 as sgm
    for n in (0, 1, 5, 12, 255, 256, 257):
        for ndev in (1, 2, 3, 4, 8):
            ranges = [sgm.shard_range(n, ndev, g) for g in range(ndev)]
            assert [k for lo, hi in ranges for k in range(lo, hi)] == list(range(n))          # every pair once, in order
            assert all(0 <= hi - lo <= -(-n // ndev) for lo, hi in ranges)                    # balanced to within one
            assert ranges == [(n * g // ndev, n * (g + 1) // ndev) for g in range(ndev)]     # documented closed form
    assert [sgm.shard_range(256, 8, g) for g in range(8)] == [(32 * g, 32 * g + 32) for g in range(8)]   # config C4 on 8 GPUs
    for bad in ((-1, 2, 0), (4, 0, 0), (4, 2, 2), (4, 2, -1)):
        with pytest.raises(sgm.SGMError):
            sgm.shard_range(*bad)


def test_two_ranks_cover_every_pair_once():
    sizes = [256, 5, 1, 0, 13]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, sizes, q)) for r in range(2)]
    for p in procs:
        p.start()
    tmax, gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tmax == 2.0
    for i, n in enumerate(sizes):
        (lo0, hi0), (lo1, hi1) = gathered[0][i], gathered[1][i]
        assert lo0 == 0 and hi0 == lo1 and hi1 == n, (n, gathered[0][i], gathered[1][i])
