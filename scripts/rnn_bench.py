#!/usr/bin/env python
"""DEC_LargeRNN throughput (BASELINE config 5: bi-GRU decoder, block_len 1000, 6 iterations, H = 100, random init).
Prints one JSON line; the CPU figure is torch.nn.GRU (what the reference executes) on a small sample."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import make_args
import turboae_b200 as T
from turboae_b200 import _lib
from oracle import turboae_oracle as O
L = int(os.environ.get("RNN_L", "1000")); B = int(os.environ.get("RNN_B", "2368"))
torch.manual_seed(0)
args = make_args(num_iteration=6, dec_num_unit=100, block_len=L, batch_size=B)
dec = T.DEC_LargeRNN(args, O.make_perm(L, 0)).cuda().eval()
if os.environ.get("RNN_PRECISION"): dec.precision = os.environ["RNN_PRECISION"]
rec = torch.randn(B, L, 3, device="cuda")
with torch.no_grad():
    dec(rec); torch.cuda.synchronize()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); y = dec(rec); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
# reference operator (torch.nn.GRU on the host cores) on a small sample: one stack, scaled to 12 stacks
Bc = 8
gru = torch.nn.GRU(7, 100, num_layers=2, batch_first=True, bidirectional=True)
x = torch.randn(Bc, L, 7)
torch.set_num_threads(os.cpu_count())
with torch.no_grad():
    gru(x); t0 = time.perf_counter(); gru(x); dt = time.perf_counter() - t0
if os.environ.get("RNN_PROFILE"):
    from torch.profiler import profile, ProfilerActivity
    with torch.no_grad(), profile(activities=[ProfilerActivity.CUDA]) as prof:
        dec(rec); torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=8, max_name_column_width=50), file=sys.stderr)
print(json.dumps({"what": "DEC_LargeRNN decode, block_len %d, num_iteration 6, H 100, precision %s" % (L, dec.precision), "batch": B, "ms": ms,
                  "codewords_per_s": B / (ms * 1e-3), "own_launches": _lib.launch_count() - n0,
                  "cpu_torch_gru_cw_per_s_est": Bc / (12 * dt), "cpu_cores": os.cpu_count(),
                  "finite": bool(torch.isfinite(y).all())}))
