"""Multi-GPU sharding of the hot path: one process per GPU, whole codewords per rank.

Decode needs no collective.  The only coupling is ENCBase.power_constraint (reference encoders.py:107-116), which
normalises by the mean / unbiased std of the WHOLE batch: ranks merge (sum x, sum x^2, count) -- three doubles --
with one all-reduce before `tae_power_norm_f32`.  The reference's nn.DataParallel wrapping of every sub-module
(reference encoders.py:343-349, decoders.py:194-199) is replaced by this; the functions work on any backend
(NCCL on the GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_codewords: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous [begin, end) of the batch owned by `rank`; sizes differ by at most one codeword."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank %d / world_size %d" % (rank, world_size))
    base, rem = divmod(n_codewords, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def merge_power_stats(stats: torch.Tensor, group=None) -> torch.Tensor:
    """In-place sum over ranks of the float64 triple (sum x, sum x^2, count) produced by `tae_enc_forward`."""
    if stats.dtype != torch.float64 or stats.numel() != 3:
        raise ValueError("power statistics must be 3 float64 values (sum, sum of squares, count)")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def mean_std_from_stats(stats: torch.Tensor) -> tuple[float, float]:
    """(mean, unbiased std) exactly as the device kernel derives them (float64 arithmetic, rounded to float32)."""
    s1, s2, n = (float(v) for v in stats.tolist())
    mean = s1 / n
    var = (s2 - n * mean * mean) / (n - 1.0)
    return float(torch.tensor(mean, dtype=torch.float32)), float(torch.tensor(max(var, 0.0) ** 0.5, dtype=torch.float32))


def max_over_ranks(value: float, device=None, group=None) -> float:
    """Timing reduction of bench.py: a multi-GPU step takes as long as its slowest rank."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
