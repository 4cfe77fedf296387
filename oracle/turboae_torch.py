"""CPU restatement of the TurboAE hot path on torch's own CPU operators (ATen / oneDNN).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  This is the *timed CPU baseline* of ``bench.py``
(``cpu_baseline`` leg and ``--impl reference``): the reference itself is PyTorch code whose CPU path is
``torch.nn.Conv1d`` + ``F.elu`` + ``torch.nn.Linear`` + advanced-index gathers, and it cannot travel to the GPU
box (``/root/reference`` does not exist there), so the same operator sequence is restated here functionally on
plain weight tensors.  It executes the same ATen kernels, with all host threads, as the reference would.

Parity status: PINNED -- tests/test_oracle.py checks it against the fixtures generated from the unmodified
reference (tests/golden/make_golden.py) and against the numpy oracle.

Each function cites the reference lines (under /root/reference) it restates.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def _w(weights, key):
    if key not in weights:
        key = key.replace(".module.", ".")
    v = weights[key]
    return v if isinstance(v, torch.Tensor) else torch.from_numpy(np.asarray(v, dtype=np.float32))


def to_torch(weights):
    return {k: _w(weights, k) for k in weights}


def _n_layers(weights, prefix):
    n = 0
    while ("%s.module.cnns.%d.weight" % (prefix, n)) in weights or ("%s.cnns.%d.weight" % (prefix, n)) in weights:
        n += 1
    return n


def same_shape_conv1d(x, weights, prefix):
    """reference cnn_utils.py:36-46: transpose, [conv1d(k, pad k//2) + ELU] x num_layer, transpose back."""
    h = torch.transpose(x, 1, 2)
    for j in range(_n_layers(weights, prefix)):
        w = _w(weights, "%s.module.cnns.%d.weight" % (prefix, j))
        b = _w(weights, "%s.module.cnns.%d.bias" % (prefix, j))
        h = F.elu(F.conv1d(h, w, b, stride=1, padding=w.shape[2] // 2))
    return torch.transpose(h, 1, 2)


def interleave(x, p):
    """reference interleavers.py:15-21."""
    return x.permute(1, 0, 2)[p].permute(1, 0, 2)


def enc_forward(u, weights, p, prefix="enc"):
    """reference encoders.py:351-377 and power_constraint :102-116 (default branch)."""
    p = torch.as_tensor(np.asarray(p), dtype=torch.long)
    x = 2.0 * u - 1.0
    outs = []
    for i, inp in ((1, x), (2, x), (3, interleave(x, p))):
        h = same_shape_conv1d(inp, weights, "%s.enc_cnn_%d" % (prefix, i))
        outs.append(F.elu(F.linear(h, _w(weights, "%s.enc_linear_%d.module.weight" % (prefix, i)),
                                   _w(weights, "%s.enc_linear_%d.module.bias" % (prefix, i)))))
    x_tx = torch.cat(outs, dim=2)
    return (x_tx - torch.mean(x_tx)) * 1.0 / torch.std(x_tx)


def dec_forward(received, weights, p, num_iteration=6, num_iter_ft=5, extrinsic=True, prefix="dec"):
    """reference decoders.py:219-269."""
    p = torch.as_tensor(np.asarray(p), dtype=torch.long)
    rp = torch.empty_like(p)
    rp[p] = torch.arange(len(p))                                      # interleavers.py:29-33
    p, rp = p.to(received.device), rp.to(received.device)
    B, L, _ = received.shape
    r_sys = received[:, :, 0].view(B, L, 1)
    r_sys_int = interleave(r_sys, p)
    r_par1 = received[:, :, 1].view(B, L, 1)
    r_par2 = received[:, :, 2].view(B, L, 1)
    prior = torch.zeros(B, L, num_iter_ft, device=received.device)
    x_plr = None
    for idx in range(num_iteration):
        last = idx == num_iteration - 1
        x_in = torch.cat([r_sys, r_par1, prior], dim=2)
        x_plr = F.linear(same_shape_conv1d(x_in, weights, "%s.dec1_cnns.%d" % (prefix, idx)),
                         _w(weights, "%s.dec1_outputs.%d.module.weight" % (prefix, idx)),
                         _w(weights, "%s.dec1_outputs.%d.module.bias" % (prefix, idx)))
        if extrinsic:
            x_plr = x_plr - prior
        x_plr_int = interleave(x_plr, p)
        x_in = torch.cat([r_sys_int, r_par2, x_plr_int], dim=2)
        x_plr = F.linear(same_shape_conv1d(x_in, weights, "%s.dec2_cnns.%d" % (prefix, idx)),
                         _w(weights, "%s.dec2_outputs.%d.module.weight" % (prefix, idx)),
                         _w(weights, "%s.dec2_outputs.%d.module.bias" % (prefix, idx)))
        if not last:
            if extrinsic:
                x_plr = x_plr - x_plr_int
            prior = interleave(x_plr, rp)                             # DeInterleaver, interleavers.py:43-48
    return torch.sigmoid(interleave(x_plr, rp))


def dec_rnn_forward(received, rnns1, rnns2, outs1, outs2, p, num_iter_ft=5, extrinsic=True):
    """DEC_LargeRNN.forward (reference decoders.py:86-149) on torch CPU operators, differentiable: `rnns*` are lists of
    torch.nn.GRU(2 + F, H, num_layers=2, batch_first=True, bidirectional=True), `outs*` lists of torch.nn.Linear -- the very
    operators the reference executes (dropout 0, dec_act 'linear')."""
    B, L, _ = received.shape
    idx = torch.as_tensor(np.asarray(p), dtype=torch.long)
    inv = torch.empty_like(idx)
    inv[idx] = torch.arange(L)
    il = lambda x: x[:, idx, :]
    dil = lambda x: x[:, inv, :]
    r_sys, r_par1, r_par2 = received[:, :, 0:1], received[:, :, 1:2], received[:, :, 2:3]
    r_sys_int = il(r_sys)
    prior = torch.zeros(B, L, num_iter_ft)
    n_it = len(rnns1)
    for i in range(n_it):
        x_plr = outs1[i](rnns1[i](torch.cat([r_sys, r_par1, prior], dim=2))[0])
        if extrinsic:
            x_plr = x_plr - prior
        x_plr_int = il(x_plr)
        x_plr = outs2[i](rnns2[i](torch.cat([r_sys_int, r_par2, x_plr_int], dim=2))[0])
        if i < n_it - 1:
            if extrinsic:
                x_plr = x_plr - x_plr_int
            prior = dil(x_plr)
    return torch.sigmoid(dil(x_plr))
