"""N > 1 host logic on CPU: world_size-2 `gloo` process group.

What is exercised is exactly what the GPU ranks run around the kernels: the contiguous codeword split, the 3-double
all-reduce that makes ENCBase.power_constraint (reference encoders.py:107-116) see the WHOLE batch, and the
max-over-ranks timing reduction of bench.py.  The per-rank kernel outputs are stood in for by the CPU oracle
(test infrastructure), so the check is: sharded statistics -> same codes as the unsharded reference computation."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ROOT, gen_inputs, load_npz


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import turboae_oracle as O
    from turboae_b200 import shard
    w = load_npz("weights_c1.npz")
    p = O.make_perm(100, 0)
    u, _ = gen_inputs(4321, B, 100, 0.0)                      # every rank draws the same full batch, keeps its shard
    lo, hi = shard.shard_range(B, rank, world)
    x_tx = O.enc_forward_unnormalised(u[lo:hi], w, p)         # stands in for tae_enc_forward on this rank's codewords
    xd = x_tx.astype(np.float64)
    stats = torch.tensor([xd.sum(), (xd * xd).sum(), float(xd.size)], dtype=torch.float64)
    shard.merge_power_stats(stats)                            # the collective under test
    mean, std = shard.mean_std_from_stats(stats)
    codes = ((x_tx - np.float32(mean)) / np.float32(std)).astype(np.float32)
    slowest = shard.max_over_ranks(10.0 + rank)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), codes=codes, lo=lo, hi=hi, mean=mean, std=std, slowest=slowest,
             count=float(stats[2]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_covers_batch_exactly():
    from turboae_b200 import shard
    for B in (0, 1, 7, 500, 50000, 50001):
        for world in (1, 2, 3, 8):
            edges = [shard.shard_range(B, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == B
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.shard_range(10, 2, 2)


def test_two_rank_power_normalisation_matches_unsharded_reference(tmp_path):
    from oracle import turboae_oracle as O
    B, world = 37, 2                                          # odd batch: ranks own 19 and 18 codewords
    mp.spawn(_worker, args=(world, _free_port(), B, str(tmp_path)), nprocs=world, join=True)
    w = load_npz("weights_c1.npz")
    u, _ = gen_inputs(4321, B, 100, 0.0)
    ref = O.enc_forward(u, w, O.make_perm(100, 0))            # the reference's batch-global normalisation
    got = np.zeros_like(ref)
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        got[int(z["lo"]):int(z["hi"])] = z["codes"]
        assert float(z["slowest"]) == 10.0 + world - 1        # MAX over ranks
        assert float(z["count"]) == B * 100 * 3               # statistics cover the WHOLE batch
    np.testing.assert_allclose(got, ref, atol=2e-6, rtol=0)
    # per-rank statistics instead would miss the tolerance: the all-reduce is what keeps parity
    local = O.enc_forward(u[:19], w, O.make_perm(100, 0))
    assert np.abs(local - ref[:19]).max() > 1e-4


def _grad_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from turboae_b200 import shard
    torch.manual_seed(5)
    full = torch.randn(8, 10, 3)
    g_full = torch.randn(8, 10, 3)
    lo, hi = shard.shard_range(8, rank, world)
    x = full[lo:hi].clone().requires_grad_(True)
    y = shard.PowerNorm.apply(x, None)                       # default group = WORLD
    y.backward(g_full[lo:hi])
    # data-parallel gradient averaging through the optimizer hook
    lin = torch.nn.Linear(4, 2)
    torch.manual_seed(100 + rank)
    lin(torch.randn(3, 4)).sum().backward()
    local = [p.grad.clone() for p in lin.parameters()]
    handle = shard.install_optimizer_hook()
    opt = torch.optim.SGD(lin.parameters(), lr=0.0)
    opt.step()
    handle.remove()
    # an UN-sharded call inside a multi-rank job (only this rank makes it, e.g. bench.py's rank-0 legs): must not enter a collective
    y_local = None
    if rank == 0:
        xl = full.clone().requires_grad_(True)
        y_local = shard.PowerNorm.apply(xl, shard.LOCAL)
        y_local.backward(g_full)
        y_local = (y_local.detach(), xl.grad.clone(),
                   shard.merge_power_stats(torch.tensor([1.0, 2.0, 3.0], dtype=torch.float64), shard.LOCAL).tolist())
    # gradients handed out as consecutive views of one flat buffer (the tensor-core training path): reduced in place, one collective
    torch.manual_seed(200 + rank)
    flat = torch.randn(2 + 12 + 5)                          # leading padding, then a (3, 4) and a (5,) view
    pa, pb = torch.nn.Parameter(torch.zeros(3, 4)), torch.nn.Parameter(torch.zeros(5))
    flat_local = flat.clone()
    pa.grad, pb.grad = flat[2:14].view(3, 4), flat[14:19]
    n_red = shard.all_reduce_gradients([pb, pa])               # any order: the views tile one range of the buffer
    torch.save({"y": y.detach(), "dx": x.grad, "lo": lo, "hi": hi, "local": local, "avg": [p.grad.clone() for p in lin.parameters()],
                "flat_local": flat_local, "flat_after": flat.clone(), "n_red": n_red, "pa": pa.grad.clone(), "y_local": y_local},
               os.path.join(out_dir, "g%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_power_norm_backward_and_gradient_all_reduce(tmp_path):
    world = 2
    mp.spawn(_grad_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    torch.manual_seed(5)
    full = torch.randn(8, 10, 3, requires_grad=True)
    g_full = torch.randn(8, 10, 3)
    y = (full - torch.mean(full)) / torch.std(full)          # reference encoders.py:107-116 on the unsharded batch
    y.backward(g_full)
    parts = [torch.load(os.path.join(str(tmp_path), "g%d.pt" % r)) for r in range(world)]
    for z in parts:
        np.testing.assert_allclose(z["y"].numpy(), y.detach()[z["lo"]:z["hi"]].numpy(), atol=1e-6)
        np.testing.assert_allclose(z["dx"].numpy(), full.grad[z["lo"]:z["hi"]].numpy(), atol=1e-6)
    for i in range(2):
        mean = (parts[0]["local"][i] + parts[1]["local"][i]) / 2
        for z in parts:
            np.testing.assert_allclose(z["avg"][i].numpy(), mean.numpy(), atol=1e-7)
    yl, dxl, st = parts[0]["y_local"]
    np.testing.assert_allclose(yl.numpy(), y.detach().numpy(), atol=1e-6)
    np.testing.assert_allclose(dxl.numpy(), full.grad.numpy(), atol=1e-6)
    assert st == [1.0, 2.0, 3.0] and parts[1]["y_local"] is None
    mean_flat = (parts[0]["flat_local"] + parts[1]["flat_local"]) / 2
    for z in parts:
        assert z["n_red"] == 17
        np.testing.assert_allclose(z["flat_after"][2:].numpy(), mean_flat[2:].numpy(), atol=1e-7)
        np.testing.assert_array_equal(z["flat_after"][:2].numpy(), z["flat_local"][:2].numpy())      # outside the views: untouched
        np.testing.assert_allclose(z["pa"].numpy(), mean_flat[2:14].view(3, 4).numpy(), atol=1e-7)


def test_single_process_is_a_no_op():
    from turboae_b200 import shard
    s = torch.tensor([1.0, 2.0, 3.0], dtype=torch.float64)
    assert shard.merge_power_stats(s.clone()).tolist() == s.tolist()
    assert shard.max_over_ranks(3.5) == 3.5
    with pytest.raises(ValueError):
        shard.merge_power_stats(torch.zeros(3))
