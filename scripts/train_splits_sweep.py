#!/usr/bin/env python
"""Graph-replayed training step (decoder mode, batch 1000) for forced numbers of group ranges per weight-gradient job family
(`dec.wgrad_splits`; None = the cost model of train_tc.wgrad_jobs)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.nn.functional as Fn
from helpers import make_args
import turboae_b200 as T
from oracle import turboae_oracle as O      # permutation helper only

dev = torch.device("cuda", 0)
B = int(os.environ.get("TRAIN_B", "1000"))
res = {}
for splits in (None, 2, 3, 4, 6, 8, None):
    torch.manual_seed(1)
    args = make_args(batch_size=B)
    p = O.make_perm(100, 0)
    enc, dec = T.ENC_interCNN(args, p).to(dev), T.DEC_LargeCNN(args, p).to(dev)
    if splits is not None:
        dec.wgrad_splits = splits
    opt = torch.optim.Adam(dec.parameters(), lr=1e-4, capturable=True, fused=True)

    def step():
        opt.zero_grad(set_to_none=True)
        u = torch.randint(0, 2, (B, 100, 1), device=dev).float()
        loss = Fn.binary_cross_entropy(torch.clamp(dec(enc(u) + torch.randn(B, 100, 3, device=dev)), 0.0, 1.0), u)
        loss.backward()
        opt.step()
        return loss.detach()
    g = T.graphs.GraphedStep(step, warmup=3, device=dev)
    for _ in range(5):
        g()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        g()
    e1.record(); torch.cuda.synchronize()
    res.setdefault(str(splits), []).append(round(e0.elapsed_time(e1) / 50, 4))
print(json.dumps(res))
