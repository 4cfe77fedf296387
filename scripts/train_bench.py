#!/usr/bin/env python
"""Training-step throughput (BASELINE config 4: enc2/dec5, batch 1000 per GPU, Adam): one `trainer.train` step =
forward (enc -> AWGN -> dec) + clamp + BCE + backward + optimizer step (reference trainer.py:41-76).  Run with python
(1 GPU) or torchrun (N GPUs, gradient all-reduce through the optimizer hook).  Prints one JSON line on rank 0."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist
import torch.nn.functional as Fn
from helpers import make_args
import turboae_b200 as T
from turboae_b200 import shard, _lib
from oracle import turboae_oracle as O      # permutation helper only

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    os.environ["TURBOAE_B200_SHARD"] = "1"
    shard.install_optimizer_hook()
B = int(os.environ.get("TRAIN_B", "1000")); steps = int(os.environ.get("TRAIN_STEPS", "8"))
torch.manual_seed(1 + rank)
args = make_args(batch_size=B)
p = O.make_perm(100, 0)
enc, dec = T.ENC_interCNN(args, p).to(dev), T.DEC_LargeCNN(args, p).to(dev)
res = {}
for mode, params in (("decoder", dec.parameters()), ("encoder", enc.parameters())):
    opt = torch.optim.Adam(params, lr=1e-4, fused=os.environ.get("TRAIN_FUSED_ADAM", "1") == "1")   # (the launcher's default)
    def step():
        opt.zero_grad()
        u = torch.randint(0, 2, (B, 100, 1), device=dev).float()
        out = dec(enc(u) + torch.randn(B, 100, 3, device=dev))
        loss = Fn.binary_cross_entropy(torch.clamp(out, 0.0, 1.0), u)
        loss.backward()
        opt.step()
        return loss
    for _ in range(2): step()
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): loss = step()
    e1.record(); torch.cuda.synchronize()
    ms = shard.max_over_ranks(e0.elapsed_time(e1) / steps, device=dev)
    res[mode] = {"ms_per_step": ms, "codewords_per_s": world * B / (ms * 1e-3), "loss": float(loss), "launches_per_step": (_lib.launch_count() - n0) / steps}
if rank == 0:
    print(json.dumps({"what": "training step (fwd+bwd+Adam), enc2/dec5, decoder train_precision=%s" % dec.train_precision, "n_gpus": world, "batch_per_gpu": B, **res}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
