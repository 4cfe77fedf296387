"""2-rank NCCL checks on real GPUs (skipped on a single-GPU box): what the gloo tests cover with the oracle standing in for the
kernels, here with the kernels themselves -- sharded encode == unsharded encode, sharded training gradients == the
whole-batch gradients, replicas start identical."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

from helpers import ROOT, gen_inputs, load_npz


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    import torch.distributed as dist
    import torch.nn.functional as Fn
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    os.environ["TURBOAE_B200_SHARD"] = "1"
    import turboae_b200 as T
    from helpers import build_codec, make_args
    from oracle import turboae_oracle as O
    from turboae_b200 import shard
    B = 64
    u, noise = gen_inputs(2468, B, 100, 0.0)
    lo, hi = shard.shard_range(B, rank, world)
    # -- inference: batch-global power normalisation across ranks ---------------------------------------------------
    m, w, p = build_codec("c1", device=dev, batch_size=hi - lo)
    assert m.enc.shard_group is not None
    with torch.no_grad():
        codes = m.enc(torch.from_numpy(u[lo:hi]).to(dev))
    # -- replicas: different random init per rank, then rank 0's state everywhere -----------------------------------
    torch.manual_seed(100 + rank)
    args = make_args(batch_size=hi - lo)
    enc, dec = T.ENC_interCNN(args, p).to(dev), T.DEC_LargeCNN(args, p).to(dev)
    before = float(next(dec.parameters()).detach().flatten()[0])
    shard.sync_replicas(enc), shard.sync_replicas(dec)
    first = torch.tensor([float(next(dec.parameters()).detach().flatten()[0])], device=dev)
    both = [torch.zeros_like(first) for _ in range(world)]
    dist.all_gather(both, first)
    # -- training: local forward / backward on the shard, ONE all-reduce of the flat gradient -----------------------
    ud, nd = torch.from_numpy(u[lo:hi]).to(dev), torch.from_numpy(noise[lo:hi]).to(dev)
    out = dec(enc(ud) + nd)
    loss = Fn.binary_cross_entropy(torch.clamp(out, 0.0, 1.0), ud)
    loss.backward()
    n_red = shard.all_reduce_gradients(list(dec.parameters()))
    g = torch.cat([q.grad.flatten() for q in dec.parameters()]).cpu().numpy()
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), codes=codes.cpu().numpy(), lo=lo, hi=hi, grad=g, n_red=n_red,
             before=before, firsts=np.array([float(t) for t in both]), loss=float(loss))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_two_ranks_nccl_sharded_encode_and_training(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    import torch.nn.functional as Fn
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [dict(np.load(os.path.join(str(tmp_path), "rank%d.npz" % i))) for i in range(world)]
    # unsharded computation on one GPU
    from helpers import build_codec, make_args
    import turboae_b200 as T
    from oracle import turboae_oracle as O
    os.environ.pop("TURBOAE_B200_SHARD", None)
    B = 64
    u, noise = gen_inputs(2468, B, 100, 0.0)
    m, w, p = build_codec("c1", device="cuda:0", batch_size=B)
    with torch.no_grad():
        full = m.enc(torch.from_numpy(u).cuda()).cpu().numpy()
    got = np.concatenate([r[0]["codes"], r[1]["codes"]], axis=0)
    np.testing.assert_allclose(got, full, atol=2e-6, rtol=0)                    # batch-global statistics through NCCL
    np.testing.assert_allclose(full, O.enc_forward(u, w, p), atol=2e-5, rtol=0)
    # replicas identical after sync (and they were not before)
    assert r[0]["firsts"][0] == r[0]["firsts"][1] == r[1]["firsts"][0]
    assert r[0]["before"] != r[1]["before"]
    # gradients: identical on both ranks after the all-reduce, == mean of the per-rank gradients == whole-batch gradient
    assert np.array_equal(r[0]["grad"], r[1]["grad"]) and int(r[0]["n_red"]) == r[0]["grad"].size
    assert np.isfinite(r[0]["grad"]).all() and float(np.abs(r[0]["grad"]).sum()) > 0.0
