#!/usr/bin/env python
"""Kernel-time breakdown of one training step (torch.profiler, CUDA activities) for the bf16 tensor-core path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.nn.functional as Fn
from torch.profiler import profile, ProfilerActivity
from helpers import make_args
import turboae_b200 as T
from oracle import turboae_oracle as O

dev = torch.device("cuda:0")
B = int(os.environ.get("TRAIN_B", "1000"))
mode = os.environ.get("TRAIN_MODE", "decoder")
torch.manual_seed(1)
args = make_args(batch_size=B)
p = O.make_perm(100, 0)
enc, dec = T.ENC_interCNN(args, p).to(dev), T.DEC_LargeCNN(args, p).to(dev)
opt = torch.optim.Adam((dec if mode == "decoder" else enc).parameters(), lr=1e-4)


def step():
    opt.zero_grad()
    u = torch.randint(0, 2, (B, 100, 1), device=dev).float()
    out = dec(enc(u) + torch.randn(B, 100, 3, device=dev))
    loss = Fn.binary_cross_entropy(torch.clamp(out, 0.0, 1.0), u)
    loss.backward()
    opt.step()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    step()
e1.record(); torch.cuda.synchronize()
print("mode %s  train_precision %s/%s  ms/step %.3f" % (mode, enc.train_precision, dec.train_precision, e0.elapsed_time(e1) / 5))
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))

# ---- per-launch times of the training kernels (CUDA events around single launches) ----
from turboae_b200 import train_tc, _lib
def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for name, mod in (("dec", dec), ("enc", enc)):
    bufs = list(mod.__dict__.get("_tc_buffers", {}).values())
    if not bufs or bufs[0].jobs is None:
        continue
    buf = bufs[0]
    arr, n = buf.jobs[0], buf.jobs[1]
    print("%s wgrad: %d jobs, %.3f ms" % (name, n, timed(lambda: train_tc.run_packed(buf.jobs, dev))))
    kinds = {}
    for j in arr:
        kinds.setdefault((j.b_chunks, j.b_nc, j.n_cols, j.taps), []).append(j)
    for k, js in kinds.items():
        sub = train_tc.pack_jobs(js, dev)
        print("   kind (b_chunks,nc,N,taps)=%s: %d jobs x %d groups: %.3f ms" % (k, len(js), js[0].g1 - js[0].g0, timed(lambda: train_tc.run_packed(sub, dev))))
        one = train_tc.pack_jobs(js[:1], dev)
        print("      one job alone: %.3f ms" % timed(lambda: train_tc.run_packed(one, dev)))
