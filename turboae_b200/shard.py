"""Multi-GPU sharding of the hot path: one process per GPU, whole codewords per rank.

Decode needs no collective.  The only coupling is ENCBase.power_constraint (reference encoders.py:107-116), which
normalises by the mean / unbiased std of the WHOLE batch: ranks merge (sum x, sum x^2, count) -- three doubles --
with one all-reduce before `tae_power_norm_f32`.  The reference's nn.DataParallel wrapping of every sub-module
(reference encoders.py:343-349, decoders.py:194-199) is replaced by this; the functions work on any backend
(NCCL on the GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist


#: pass as `group` to keep a reduction local to this rank even though torch.distributed is initialised (an un-sharded module
#: inside a multi-rank job: only some ranks call it, so it must not enter a collective)
LOCAL = False


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
    if group is LOCAL:
        return stats
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


class PowerNorm(torch.autograd.Function):
    """ENCBase.power_constraint (reference encoders.py:107-116), differentiable and exact under sharding:
    y = (x - mean) / std with mean / unbiased std over the WHOLE batch.  Forward all-reduces (sum x, sum x^2, count),
    backward all-reduces (sum g, sum g*y):  dx = (g - mean(g) - y * sum(g*y) / (N - 1)) / std.

    On a CUDA device both directions are this package's kernels (``tae_power_stats_f32`` unless the encoder kernel already
    delivered ``own_stats``, ``tae_power_norm_f32``, ``tae_power_norm_bwd_sums_f32``, ``tae_power_norm_bwd_f32``): three launches
    forward and backward together instead of ~25 small torch operators.  ``own_stats`` (3 device doubles: this rank's sum, sum of
    squares, count) is merged across the ranks IN PLACE and kept for the backward: pass a tensor nobody else reads afterwards.  The
    torch spelling below serves the host-side tests of the collective logic (gloo, CPU tensors)."""

    @staticmethod
    def forward(ctx, x, group, own_stats=None):
        ctx.group = group
        ctx.native = x.is_cuda and x.dtype == torch.float32
        if ctx.native:
            from . import _lib
            lib = _lib.load()
            x = x.contiguous()
            with torch.cuda.device(x.device):
                stream = _lib.stream_ptr(x.device)
                stats = own_stats
                if stats is None:
                    stats = torch.zeros(3, dtype=torch.float64, device=x.device)
                    _lib.check(lib.tae_power_stats_f32(_lib.ptr(x), x.numel(), _lib.ptr(stats), stream))
                merge_power_stats(stats, group)
                y = torch.empty_like(x)
                mean_std = torch.empty(2, dtype=torch.float32, device=x.device)
                _lib.check(lib.tae_power_norm_f32(_lib.ptr(x), _lib.ptr(y), x.numel(), _lib.ptr(stats), _lib.ptr(mean_std), stream))
            ctx.save_for_backward(y, stats, mean_std)
            return y
        xd = x.double()
        # (torch.full, not torch.tensor: a host scalar copied to the device would make the host wait for the stream)
        stats = torch.stack([xd.sum(), (xd * xd).sum(), torch.full((), float(x.numel()), dtype=torch.float64, device=x.device)])
        merge_power_stats(stats, group)
        n = stats[2]
        mean = stats[0] / n
        std = torch.sqrt(torch.clamp((stats[1] - n * mean * mean) / (n - 1.0), min=0.0))
        y = (x - mean.float()) / std.float()
        ctx.save_for_backward(y, std.float(), n)
        return y

    @staticmethod
    def backward(ctx, g):
        sharded = ctx.group is not LOCAL and dist.is_available() and dist.is_initialized() and dist.get_world_size(ctx.group) > 1
        if ctx.native:
            from . import _lib
            lib = _lib.load()
            y, stats, mean_std = ctx.saved_tensors
            g = g.to(torch.float32).contiguous()
            with torch.cuda.device(g.device):
                stream = _lib.stream_ptr(g.device)
                sums = torch.zeros(2, dtype=torch.float64, device=g.device)
                _lib.check(lib.tae_power_norm_bwd_sums_f32(_lib.ptr(g), _lib.ptr(y), g.numel(), _lib.ptr(sums), stream))
                if sharded:
                    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=ctx.group)
                dx = torch.empty_like(g)
                _lib.check(lib.tae_power_norm_bwd_f32(_lib.ptr(g), _lib.ptr(y), _lib.ptr(dx), g.numel(), _lib.ptr(sums), _lib.ptr(stats),
                                                      _lib.ptr(mean_std), stream))
            return dx, None, None
        y, std, n = ctx.saved_tensors
        sums = torch.stack([g.double().sum(), (g.double() * y.double()).sum()])
        if sharded:
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=ctx.group)
        dx = (g - (sums[0] / n).float() - y * (sums[1] / (n - 1.0)).float()) / std
        return dx, None, None


def seed_everything(seed: int) -> None:
    """python / numpy / torch (CPU and CUDA) generators from one integer (the reference seeds nothing: main.py)."""
    import random
    import numpy as np
    seed = int(seed) % (2 ** 32)
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


def sync_replicas(module: torch.nn.Module, src: int = 0, group=None) -> int:
    """Broadcast every parameter and buffer of `module` from rank `src`: data-parallel replicas must start identical
    (the reference's nn.DataParallel replicates one model; one process per GPU has to do it explicitly).  Returns the number
    of tensors sent.  A no-op outside a multi-rank job."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0
    n = 0
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src, group=group)
            n += 1
    from . import _flat
    _flat.invalidate_all()          # written through .data: no version counter moved
    return n


def all_reduce_gradients(params, group=None) -> int:
    """Data-parallel training (BASELINE config 4): ONE all-reduce (average) of the flat gradient of `params` -- the
    replacement of nn.DataParallel's ReduceAddCoalesced.  Returns the number of floats reduced."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return sum(g.numel() for g in grads)
    # The tensor-core training path hands out gradients as consecutive views of ONE flat buffer (train_tc.py): reduce it in
    # place -- one collective, no gather / scatter copies.
    base = grads[0].untyped_storage()
    if all(g.is_contiguous() and g.untyped_storage().data_ptr() == base.data_ptr() for g in grads):
        spans = sorted((g.storage_offset(), g.numel()) for g in grads)      # any parameter order: the views must tile one range
        offs = [o for o, _ in spans]
        if all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(len(spans) - 1)):
            n = spans[-1][0] + spans[-1][1] - offs[0]
            flat = torch.empty(0, dtype=grads[0].dtype, device=grads[0].device).set_(base, offs[0], (n,), (1,))
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
            flat /= dist.get_world_size(group)
            return n
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat /= dist.get_world_size(group)
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()
    return off


def install_optimizer_hook(group=None):
    """Registers a global optimizer pre-step hook that all-reduces the gradients of whatever the optimizer is about to
    step: the reference's trainer.py (`loss.backward(); optimizer.step()`, :74-76) then trains data-parallel unchanged."""
    from torch.optim.optimizer import register_optimizer_step_pre_hook

    def hook(optimizer, args, kwargs):
        params = [p for grp in optimizer.param_groups for p in grp["params"]]
        all_reduce_gradients(params, group)

    return register_optimizer_step_pre_hook(hook)
