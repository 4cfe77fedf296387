"""Training of DEC_LargeCNN on the tensor cores (reference trainer.py:33-76 backpropagating
through decoders.py:219-269 and cnn_utils.py:36-46).

One ``torch.autograd.Function`` for the whole decoder:

* forward  = the fused inference kernel (``tae_dec_forward_train_bf16``: same launch, same schedule) that additionally
  stashes every conv layer's output and every stack's input as bf16 *group images* in HBM;
* backward = per stack, in reverse order, ``tae_dec_stack_backward_bf16`` (the same tcgen05 pipeline run on gradients,
  ELU' taken from the stash) with the extrinsic / interleaver glue between stacks spelled out on (B, L, F) tensors,
  then ONE ``tae_wgrad_bf16`` launch that turns the stashed gradients and activations into all weight gradients.

Operands are bf16, every accumulation (MMA, weight-gradient reduction) is fp32, parameters and their gradients stay fp32.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._flat import unwrap

N_SM = 148


def _odd_chunks(n_ch: int, c0: int) -> int:
    """chunks of 8 channels to read from c0 so that a spare (bias) column follows: an odd count (8*nc + 8 = UMMA N % 16)."""
    nc = (n_ch + 7) // 8
    return nc if nc % 2 == 1 else nc + 1


class _Buffers:
    """Zero-initialised group-image buffers of one (batch, block length) shape, reused across steps."""

    def __init__(self, n_stacks, n_layer, groups, device):
        cb = _lib.IMG_CHUNK_BYTES
        z = lambda n: torch.zeros(n, dtype=torch.uint8, device=device)
        self.groups = groups
        self.stash_y = z(n_stacks * n_layer * groups * _lib.IMG_CHUNKS * cb)
        self.stash_g = z(n_stacks * n_layer * groups * _lib.IMG_CHUNKS * cb)
        self.stash_x = z(n_stacks * groups * cb)
        self.stash_d = z(n_stacks * groups * cb)


def _buffers(dec, n_stacks, n_layer, groups, device):
    key = (n_stacks, n_layer, groups, str(device))
    cache = dec.__dict__.setdefault("_tc_buffers", {})
    if key not in cache:
        cache.clear()                      # one live shape at a time: the images are large
        cache[key] = _Buffers(n_stacks, n_layer, groups, device)
    return cache[key]


def wgrad_jobs(n_layer, units, cin0, fouts, groups, stash_y, stash_x, stash_g, stash_d, gflat, offsets, splits=1):
    """Job list of ``tae_wgrad_bf16`` for conv stacks laid out like a DEC_LargeCNN: ``offsets[st]`` = (per layer (w_off, b_off),
    lin_w_off) in floats into ``gflat``; ``fouts[st]`` = features of the stack's Linear."""
    cb = _lib.IMG_CHUNK_BYTES
    layer_img = groups * _lib.IMG_CHUNKS * cb
    jobs = []

    def add(a, b, grad, bias, b_chunks, c0, nc, taps, n_cols, m_valid, n_valid, n0, s_m, s_n, s_t):
        for sp in range(splits):
            g0, g1 = groups * sp // splits, groups * (sp + 1) // splits
            if g1 > g0:
                jobs.append(_lib.TaeWgradJob(a, b, grad, bias, b_chunks, c0, nc, taps, n_cols, m_valid, n_valid, n0,
                                             s_m, s_n, s_t, g0, g1, 0))

    gp = gflat.data_ptr()
    for st, (layers, lin_w_off) in enumerate(offsets):
        y = stash_y.data_ptr() + st * n_layer * layer_img
        g = stash_g.data_ptr() + st * n_layer * layer_img
        x = stash_x.data_ptr() + st * groups * cb
        d = stash_d.data_ptr() + st * groups * cb
        for j in range(n_layer - 1, 0, -1):          # the big jobs first
            w_off, b_off = layers[j]
            a_img, b_img = g + j * layer_img, y + (j - 1) * layer_img
            c0 = 0
            while c0 * 8 < units:
                rest = units - c0 * 8
                if rest > 64:                        # a full 8-chunk slab, no spare column
                    add(a_img, b_img, gp + 4 * w_off, None, _lib.IMG_CHUNKS, c0, 8, 5, 64, units, 64, c0 * 8, 5 * units, 5, 1)
                    c0 += 8
                elif rest > 56:                      # 7 chunks; the bias column comes with the next (last) slab
                    add(a_img, b_img, gp + 4 * w_off, None, _lib.IMG_CHUNKS, c0, 7, 5, 64, units, 56, c0 * 8, 5 * units, 5, 1)
                    c0 += 7
                else:
                    nc = _odd_chunks(rest, c0)
                    add(a_img, b_img, gp + 4 * w_off, gp + 4 * b_off, _lib.IMG_CHUNKS, c0, nc, 5, 8 * nc + 8, units, rest, c0 * 8,
                        5 * units, 5, 1)
                    c0 += nc
        w_off, b_off = layers[0]
        add(g, x, gp + 4 * w_off, gp + 4 * b_off, 1, 0, 1, 5, 16, units, cin0, 0, 5 * cin0, 5, 1)
        add(y + (n_layer - 1) * layer_img, d, gp + 4 * lin_w_off, None, 1, 0, 1, 1, 16, units, fouts[st], 0, 1, units, 0)
    return jobs


def run_wgrad(jobs, device):
    lib = _lib.load()
    arr = (_lib.TaeWgradJob * len(jobs))(*jobs)
    ws = torch.empty(256 + C.sizeof(_lib.TaeWgradJob) * len(jobs) + 64, dtype=torch.uint8, device=device)
    _lib.check(lib.tae_wgrad_bf16(arr, len(jobs), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(device)))
    return ws


class DecoderTrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dec, received, *params):
        lib = _lib.load()
        a = dec.args
        B, L, _ = received.shape
        dev = received.device
        n_stacks, n_layer = 2 * a.num_iteration, a.dec_num_layer
        with torch.cuda.device(dev):
            _, cfg, flat, packed, _ = dec._prepare(L, dev, "bf16")
            groups = lib.tae_train_groups(L, B)
            buf = _buffers(dec, n_stacks, n_layer, groups, dev)
            perm, inv = dec.interleaver.device_index(dev)
            out = torch.empty((B, L, 1), dtype=torch.float32, device=dev)
            ws = dec._ws.get(256, dev)
            _lib.check(lib.tae_dec_forward_train_bf16(cfg, _lib.ptr(packed), _lib.ptr(received), _lib.ptr(perm), _lib.ptr(inv),
                                                      _lib.ptr(out), None, B, _lib.ptr(buf.stash_y), _lib.ptr(buf.stash_x),
                                                      _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)))
        ctx.dec, ctx.cfg, ctx.buf = dec, cfg, buf
        ctx.shape = (B, L)
        ctx.need_input = received.requires_grad
        ctx.need_params = any(p.requires_grad for p in params)
        ctx.save_for_backward(out, flat)
        return out

    @staticmethod
    def backward(ctx, d_out):
        lib = _lib.load()
        dec, cfg, buf = ctx.dec, ctx.cfg, ctx.buf
        a = dec.args
        out, flat = ctx.saved_tensors
        B, L = ctx.shape
        dev = d_out.device
        F, units, n_layer, I = a.num_iter_ft, a.dec_num_unit, a.dec_num_layer, a.num_iteration
        n_stacks = 2 * I
        with torch.cuda.device(dev):
            packed_bwd = dec._flat.derived.get("bf16_bwd")
            if packed_bwd is None:
                packed_bwd = torch.empty(lib.tae_dec_bwd_packed_bytes(cfg), dtype=torch.uint8, device=dev)
                _lib.check(lib.tae_dec_pack_bwd_bf16(cfg, _lib.ptr(flat), _lib.ptr(packed_bwd), _lib.stream_ptr(dev)))
                dec._flat.derived["bf16_bwd"] = packed_bwd
            perm, inv = dec.interleaver.device_index(dev)
            perm_l, inv_l = perm.long(), inv.long()
            ws = dec._ws.get(256, dev)
            d_rec = torch.zeros((B, L, 3), dtype=torch.float32, device=dev)
            gflat = torch.zeros_like(flat)
            # out = sigmoid(deinterleave(o_last))  (decoders.py:267)  =>  d o_last = interleave(d_out * out * (1 - out))
            d_o = (d_out.to(torch.float32) * out * (1.0 - out)).index_select(1, perm_l).contiguous()
            dxin = torch.empty((B, L, 8), dtype=torch.float32, device=dev)
            lin_bias_grads = []
            for st in range(n_stacks - 1, -1, -1):
                fin = d_o.shape[2]
                lin_bias_grads.append((st, d_o.sum(dim=(0, 1))))
                _lib.check(lib.tae_dec_stack_backward_bf16(cfg, _lib.ptr(packed_bwd), st, _lib.ptr(d_o), fin, _lib.ptr(buf.stash_y),
                                                           _lib.ptr(buf.stash_g), _lib.ptr(buf.stash_d), _lib.ptr(dxin), B,
                                                           _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)))
                if st % 2 == 0:      # [r_sys, r_par1, prior]                                 (decoders.py:230)
                    d_rec[:, :, 0] += dxin[:, :, 0]
                    d_rec[:, :, 1] += dxin[:, :, 1]
                else:                # [interleave(r_sys), r_par2, interleave(x_plr)]          (decoders.py:240)
                    d_rec[:, :, 0] += dxin[:, :, 0].index_select(1, inv_l)
                    d_rec[:, :, 2] += dxin[:, :, 1]
                if st == 0:
                    break
                d_prior = dxin[:, :, 2:2 + F]
                if a.extrinsic and st != n_stacks - 1:
                    d_prior = d_prior - d_o            # x_plr = Linear(...) - prior            (decoders.py:235-236, 246-247)
                # the prior of stack st is interleave (st odd) / deinterleave (st even) of the previous stack's extrinsic output
                d_o = d_prior.index_select(1, inv_l if st % 2 == 1 else perm_l).contiguous()
            grads = [None] * len(dec.ordered_parameters())
            if ctx.need_params:
                offsets, fouts, off = [], [], 0
                for idx in range(I):
                    for s in range(2):
                        layers = []
                        for j in range(n_layer):
                            cin = (2 + F) if j == 0 else units
                            layers.append((off, off + units * cin * 5))
                            off += units * cin * 5 + units
                        fout = 1 if (s == 1 and idx == I - 1) else F
                        offsets.append((layers, off))
                        fouts.append(fout)
                        off += fout * units + fout
                jobs = wgrad_jobs(n_layer, units, 2 + F, fouts, buf.groups, buf.stash_y, buf.stash_x, buf.stash_g, buf.stash_d,
                                  gflat, offsets, splits=getattr(dec, "wgrad_splits", 1))
                keep = run_wgrad(jobs, dev)
                for st, gb in lin_bias_grads:
                    lin_w_off = offsets[st][1]
                    gflat[lin_w_off + fouts[st] * units: lin_w_off + fouts[st] * units + fouts[st]] = gb
                off = 0
                for i, p in enumerate(dec.ordered_parameters()):
                    n = p.numel()
                    if p.requires_grad:
                        grads[i] = gflat[off:off + n].view_as(p)
                    off += n
                del keep
        return (None, d_rec if ctx.need_input else None, *grads)


def decoder_forward_train(dec, received):
    """Differentiable DEC_LargeCNN.forward on the tensor cores."""
    if dec.args.dec_kernel_size != 5:
        raise NotImplementedError("tensor-core training needs dec_kernel_size == 5")
    params = dec.ordered_parameters()
    return DecoderTrainFn.apply(dec, received.contiguous(), *params)


class EncoderTrainFn(torch.autograd.Function):
    """ENC_interCNN branches (reference encoders.py:362-373: three conv stacks, Linear(units, 1), ELU, concat) on the tensor
    cores: u (B, L, 1) -> un-normalised x_tx (B, L, 3).  power_constraint stays differentiable torch glue (shard.PowerNorm)."""

    @staticmethod
    def forward(ctx, enc, u, *params):
        lib = _lib.load()
        a = enc.args
        B, L, _ = u.shape
        dev = u.device
        n_layer = a.enc_num_layer
        with torch.cuda.device(dev):
            cfg = enc.config(L)
            flat = enc._flat.get(enc.ordered_parameters())
            packed = enc._flat.derived.get("bf16")
            if packed is None:
                nbytes = lib.tae_enc_packed_bytes(cfg)
                if nbytes == 0:
                    raise _lib.TaeError("bf16 encoder path unavailable for this configuration (%s)" % lib.tae_last_error().decode())
                packed = torch.empty(nbytes, dtype=torch.uint8, device=dev)
                _lib.check(lib.tae_enc_pack_bf16(cfg, _lib.ptr(flat), _lib.ptr(packed), _lib.stream_ptr(dev)))
                enc._flat.derived["bf16"] = packed
            groups = lib.tae_train_groups(L, B)
            buf = _buffers(enc, 3, n_layer, groups, dev)
            perm, inv = enc.interleaver.device_index(dev)
            x_tx = torch.empty((B, L, 3), dtype=torch.float32, device=dev)
            stats = torch.zeros(3, dtype=torch.float64, device=dev)
            ws = enc._ws.get(256, dev)
            _lib.check(lib.tae_enc_forward_train_bf16(cfg, _lib.ptr(packed), _lib.ptr(u), _lib.ptr(perm), _lib.ptr(inv), _lib.ptr(x_tx),
                                                      _lib.ptr(stats), B, _lib.ptr(buf.stash_y), _lib.ptr(buf.stash_x), _lib.ptr(ws),
                                                      ws.numel(), _lib.stream_ptr(dev)))
        ctx.enc, ctx.cfg, ctx.buf, ctx.shape = enc, cfg, buf, (B, L)
        ctx.save_for_backward(x_tx, flat)
        return x_tx

    @staticmethod
    def backward(ctx, d_x):
        lib = _lib.load()
        enc, cfg, buf = ctx.enc, ctx.cfg, ctx.buf
        a = enc.args
        x_tx, flat = ctx.saved_tensors
        B, L = ctx.shape
        dev = d_x.device
        units, n_layer = a.enc_num_unit, a.enc_num_layer
        with torch.cuda.device(dev):
            packed_bwd = enc._flat.derived.get("bf16_bwd")
            if packed_bwd is None:
                packed_bwd = torch.empty(lib.tae_enc_bwd_packed_bytes(cfg), dtype=torch.uint8, device=dev)
                _lib.check(lib.tae_enc_pack_bwd_bf16(cfg, _lib.ptr(flat), _lib.ptr(packed_bwd), _lib.stream_ptr(dev)))
                enc._flat.derived["bf16_bwd"] = packed_bwd
            ws = enc._ws.get(256, dev)
            gflat = torch.zeros_like(flat)
            # x_tx = ELU(Linear(h))  =>  d lin = d x_tx * ELU'   (ELU' = x_tx + 1 where x_tx < 0)
            d_lin = (d_x.to(torch.float32) * torch.where(x_tx > 0, torch.ones_like(x_tx), x_tx + 1.0))
            dxin = torch.empty((B, L, 8), dtype=torch.float32, device=dev)
            offsets, off = [], 0
            for br in range(3):
                layers = []
                for j in range(n_layer):
                    cin = 1 if j == 0 else units
                    layers.append((off, off + units * cin * 5))
                    off += units * cin * 5 + units
                offsets.append((layers, off))
                off += units + 1
            for br in range(3):
                d_br = d_lin[:, :, br:br + 1].contiguous()
                _lib.check(lib.tae_enc_stack_backward_bf16(cfg, _lib.ptr(packed_bwd), br, _lib.ptr(d_br), _lib.ptr(buf.stash_y),
                                                           _lib.ptr(buf.stash_g), _lib.ptr(buf.stash_d), _lib.ptr(dxin), B, _lib.ptr(ws),
                                                           ws.numel(), _lib.stream_ptr(dev)))
                gflat[offsets[br][1] + units] = d_br.sum()
            jobs = wgrad_jobs(n_layer, units, 1, [1, 1, 1], buf.groups, buf.stash_y, buf.stash_x, buf.stash_g, buf.stash_d, gflat,
                              offsets, splits=getattr(enc, "wgrad_splits", 1))
            keep = run_wgrad(jobs, dev)
            grads, off = [], 0
            for p in enc.ordered_parameters():
                n = p.numel()
                grads.append(gflat[off:off + n].view_as(p) if p.requires_grad else None)
                off += n
            del keep
        return (None, None, *grads)


def encoder_branches_train(enc, u):
    """Differentiable un-normalised ENC_interCNN output (B, L, 3) on the tensor cores."""
    if enc.args.enc_kernel_size != 5:
        raise NotImplementedError("tensor-core training needs enc_kernel_size == 5")
    return EncoderTrainFn.apply(enc, u.contiguous(), *enc.ordered_parameters())
