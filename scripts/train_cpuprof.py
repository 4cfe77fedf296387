#!/usr/bin/env python
"""Host-side (Python) cost of one training step: cProfile over 10 steps."""
import cProfile, os, pstats, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.nn.functional as Fn
from helpers import make_args
import turboae_b200 as T
from oracle import turboae_oracle as O
dev = torch.device("cuda:0")
B = 1000
mode = os.environ.get("TRAIN_MODE", "decoder")
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
for _ in range(3): step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(10): step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
