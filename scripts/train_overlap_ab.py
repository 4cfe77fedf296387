#!/usr/bin/env python
"""A/B of the two-launch backward (train_tc.backward_split) on the decoder-mode training step, as a CUDA-graph replay (no host
in the timed region) and eager: ms per step with the weight gradients of the first wave beside the last wave vs one launch."""
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
res = {"batch": B}
for overlap in (False, True, False, True):
    torch.manual_seed(1)
    args = make_args(batch_size=B)
    p = O.make_perm(100, 0)
    enc, dec = T.ENC_interCNN(args, p).to(dev), T.DEC_LargeCNN(args, p).to(dev)
    dec.wgrad_overlap = overlap
    opt = torch.optim.Adam(dec.parameters(), lr=1e-4, capturable=True, fused=True)

    def step():
        opt.zero_grad(set_to_none=True)
        u = torch.randint(0, 2, (B, 100, 1), device=dev).float()
        out = dec(enc(u) + torch.randn(B, 100, 3, device=dev))
        loss = Fn.binary_cross_entropy(torch.clamp(out, 0.0, 1.0), u)
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
        loss = g()
    e1.record(); torch.cuda.synchronize()
    res.setdefault("graphed_ms_overlap_%s" % overlap, []).append(round(e0.elapsed_time(e1) / 50, 4))
    res["loss_overlap_%s" % overlap] = float(loss)
print(json.dumps(res))
