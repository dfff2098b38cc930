"""Batch sharding rule shared by the multi-process benchmark and SGMB_MatchBatchMultiGPU (sgm_b200.cu):
pair k of n goes to worker k*world//n, i.e. contiguous shards [n*r//world, n*(r+1)//world).  Independent
stereo pairs are the only unit that can be split: every aggregation path spans the whole image
(SURVEY.md section 8e), so there is no collective on the data path."""
from __future__ import annotations


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return (n * rank) // world, (n * (rank + 1)) // world
