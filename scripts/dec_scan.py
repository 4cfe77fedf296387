#!/usr/bin/env python
"""Timing scan of the fused decoder over num_layer / num_iteration (random weights): separates the per-conv-layer
cost from the per-stack overhead.  Run on the GPU box."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import make_args
import turboae_b200 as T
from oracle import turboae_oracle as O

def time_cfg(n_layer, n_iter, B=50000, reps=5):
    torch.manual_seed(0)
    args = make_args(dec_num_layer=n_layer, num_iteration=n_iter, batch_size=B)
    dec = T.DEC_LargeCNN(args, O.make_perm(100, 0)).cuda().eval()
    rec = torch.randn(B, 100, 3, device="cuda")
    with torch.no_grad():
        for _ in range(2): dec(rec)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): dec(rec)
        e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

res = {}
import ast
CFGS = ast.literal_eval(os.environ.get('SCAN', '[(2, 6), (3, 6), (5, 6), (7, 6), (5, 3), (5, 1)]'))
for n_layer, n_iter in CFGS:
    ms = time_cfg(n_layer, n_iter)
    res["L%d_I%d" % (n_layer, n_iter)] = ms
    print("num_layer %d num_iteration %d : %.3f ms" % (n_layer, n_iter, ms), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "dec_scan.json"), "w"))
