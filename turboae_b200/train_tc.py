"""Training of DEC_LargeCNN on the tensor cores (reference trainer.py:33-76 backpropagating
through decoders.py:219-269 and cnn_utils.py:36-46).

One ``torch.autograd.Function`` for the whole decoder:

* forward  = the fused inference kernel (``tae_dec_forward_train_bf16``: same launch, same schedule) that additionally
  stashes every conv layer's output and every stack's input as bf16 *group images* in HBM;
* backward = ``tae_dec_backward_bf16`` / ``tae_dec_backward_range_bf16``: all stacks in ONE launch, the schedule walked backwards
  (the same tcgen05 pipeline run on gradients, ELU' taken from the stash, the extrinsic / interleaver glue between stacks inside
  the kernel), then ``tae_wgrad_bf16`` turns the stashed gradients and activations into all weight gradients.  When the batch's
  work units fill one wave of CTA pairs plus a partly filled one (batch 1000 on a B200) the backward is split at the full wave and
  the first part's weight gradients run on a side stream beside the rest (``backward_split``).

Operands are bf16, every accumulation (MMA, weight-gradient reduction) is fp32, parameters and their gradients stay fp32.
"""
from __future__ import annotations

import ctypes as C
import weakref

import torch

import os

from . import _lib
from ._flat import unwrap

#: split the decoder's backward at the last full wave of work units and run the first part's weight gradients beside the rest
#: (backward_split); a module attribute ``wgrad_overlap`` or TURBOAE_B200_WGRAD_OVERLAP=0 turns it off
WGRAD_OVERLAP = os.environ.get("TURBOAE_B200_WGRAD_OVERLAP", "1") != "0"


def n_sm(device=None) -> int:
    """SM count of the device the job list is balanced for (148 on a B200; that figure is also used when the job list is built
    on a host without a GPU, e.g. by the CPU tests of the job construction)."""
    if not torch.cuda.is_available():
        return 148
    return torch.cuda.get_device_properties(device if device is not None else torch.cuda.current_device()).multi_processor_count


class _Token:
    """Lives exactly as long as the autograd graph node (ctx) that owns a stash: lets a later forward see whether the previous
    one is still waiting for its backward."""
    __slots__ = ("done", "__weakref__")

    def __init__(self):
        self.done = False


def _odd_chunks(n_ch: int, c0: int) -> int:
    """chunks of 8 channels to read from c0 so that a spare (bias) column follows: an odd count (8*nc + 8 = UMMA N % 16)."""
    nc = (n_ch + 7) // 8
    return nc if nc % 2 == 1 else nc + 1


class _Buffers:
    """Zero-initialised group-image buffers (and the small fp32 chain buffers) of one (batch, block length) shape, reused
    across steps; also caches everything whose device pointers are stable: the flat gradient buffer, the weight-gradient
    job array."""

    def __init__(self, n_stacks, n_layer, groups, B, L, F, device):
        cb = _lib.IMG_CHUNK_BYTES
        z = lambda n: torch.zeros(n, dtype=torch.uint8, device=device)
        self.groups = groups
        self.stash_y = z(n_stacks * n_layer * groups * _lib.IMG_CHUNKS * cb)
        self.stash_g = z(n_stacks * n_layer * groups * _lib.IMG_CHUNKS * cb)
        self.stash_x = z(n_stacks * groups * cb)
        self.stash_d = z(n_stacks * groups * cb)
        self.dxin = torch.zeros((n_stacks, B, L, 8), dtype=torch.float32, device=device)
        self.dlin = torch.zeros((n_stacks, B, L, F), dtype=torch.float32, device=device)
        self.gflat = None
        self.jobs = None          # (ctypes array, n, device copy, device workspace): all groups, or those of the first backward launch
        self.jobs_tail = None     # ... of the second backward launch (backward_split)
        self.jobs_cut = None      # the split point the two lists were built for
        self.side = None          # side stream of the overlapped weight-gradient launch
        self.gen = 0              # forwards that have written this stash
        self.owner = None         # weakref to the _Token of the forward whose backward has not run yet

    def busy(self) -> bool:
        tok = self.owner() if self.owner is not None else None
        return tok is not None and not tok.done

    def claim(self):
        """Called by a forward that is about to overwrite the stash; returns (token, generation) for its ctx."""
        tok = _Token()
        self.owner = weakref.ref(tok)
        self.gen += 1
        return tok, self.gen


def _buffers(mod, n_stacks, n_layer, groups, B, L, F, device):
    """The module's cached stash for this shape -- or a fresh, un-cached one when the cached stash still belongs to a forward
    whose backward has not run (gradient accumulation over micro-batches with one summed loss, two SNRs in one loss, ...):
    overwriting it would make that graph's backward silently wrong."""
    key = (n_stacks, n_layer, groups, B, L, F, str(device))
    cache = mod.__dict__.setdefault("_tc_buffers", {})
    if key not in cache:
        cache.clear()                      # one live shape at a time: the images are large
        cache[key] = _Buffers(n_stacks, n_layer, groups, B, L, F, device)
    buf = cache[key]
    if buf.busy():
        buf = _Buffers(n_stacks, n_layer, groups, B, L, F, device)
    return buf


def _check_stash(ctx):
    if ctx.buf.gen != ctx.gen:
        raise _lib.TaeError("the activation stash of this forward was overwritten by a later forward of the same module before "
                            "backward() ran (a graph kept alive with retain_graph, or a second backward)")


def _grad_views(gflat, params):
    """Per-parameter gradients as consecutive views of the flat buffer.  Fresh view objects on every call: autograd adopts an
    incoming gradient as ``.grad`` (no copy) only when nothing else references it, and shard.all_reduce_gradients relies on the
    ``.grad``s tiling one buffer.  One split call + one reshape per multi-dimensional parameter."""
    return [(v if p.dim() == 1 else v.view(p.shape)) if p.requires_grad else None
            for v, p in zip(gflat.split_with_sizes([p.numel() for p in params]), params)]


def _flat_grad(buf, flat, params):
    """The persistent flat gradient buffer, zeroed -- or a fresh one when some .grad still aliases it (gradient accumulation
    without zero_grad: the views handed out by the previous backward were adopted by autograd as the .grad tensors)."""
    g = buf.gflat
    if g is not None and g.numel() == flat.numel():
        base = g.untyped_storage().data_ptr()
        if not any(p.grad is not None and p.grad.untyped_storage().data_ptr() == base for p in params):
            g.zero_()
            return g, False
    buf.gflat = torch.zeros_like(flat)
    buf.jobs = None
    return buf.gflat, True


def _job_cost_us(b_chunks, n_cols, taps):
    """Measured cost of one group in one job (B200, microseconds): the MMAs are operand-fetch bound."""
    if b_chunks > 1:
        if n_cols > 64:                      # tap-split job over all input channels (N = 112): ~69 cycles per MMA; the 2-tap job waits
            return 3.2 + 0.5 * (taps - 2)    # for its loads a fifth of the time (profiles/r02_wgrad_probe.log: 3.7 / 3.2 us per group)
        return 5.0 if n_cols > 48 else 4.4
    return 3.9 if taps > 1 else 1.9


#: layers wider than 64 input channels: jobs split the TAPS ({0,1,2} and {3,4}) over all channels instead of the channels (64 + rest)
#: over all taps -- fewer operand bytes per tensor-core cycle (tae_wgrad.cu); TURBOAE_B200_WGRAD_TAPSPLIT=0 keeps the channel slabs
WGRAD_TAP_SPLIT = os.environ.get("TURBOAE_B200_WGRAD_TAPSPLIT", "1") != "0"


def wgrad_jobs(n_layer, units, cin0, fouts, groups, stash_y, stash_x, stash_g, stash_d, gflat, offsets, splits=None, sm_count=None,
               group_range=None):
    """Job list of ``tae_wgrad_bf16`` for conv stacks laid out like a DEC_LargeCNN: ``offsets[st]`` = (per layer (w_off, b_off),
    lin_w_off) in floats into ``gflat``; ``fouts[st]`` = features of the stack's Linear.  One job = one CTA.  Every
    layer is cut into group ranges of about equal estimated duration (3-4 CTAs per SM in total, at least ~60 us each so that
    the TMEM drain stays a small share), the channel slabs of a layer share the ranges and are adjacent in the list (they read
    the same gradient image: the second reader hits L2), and the list is sorted longest first: the hardware dispatches CTAs
    in order, which then balances the SMs.  ``splits`` forces the number of ranges instead.  ``group_range`` = (lo, hi) restricts
    the reduction to the groups lo .. hi-1 (the gradients are sums over groups and the kernel ADDS into ``gflat``, so the job lists
    of disjoint ranges may run as separate launches, e.g. beside a backward that is still producing the later groups)."""
    cb = _lib.IMG_CHUNK_BYTES
    layer_img = groups * _lib.IMG_CHUNKS * cb
    protos = []           # (fields, family): the channel slabs of ONE layer form a family
    family = [0]

    def add(*f):
        protos.append((f, family[0]))

    gp = gflat.data_ptr()
    for st, (layers, lin_w_off) in enumerate(offsets):
        y = stash_y.data_ptr() + st * n_layer * layer_img
        g = stash_g.data_ptr() + st * n_layer * layer_img
        x = stash_x.data_ptr() + st * groups * cb
        d = stash_d.data_ptr() + st * groups * cb
        for j in range(n_layer - 1, 0, -1):
            w_off, b_off = layers[j]
            a_img, b_img = g + j * layer_img, y + (j - 1) * layer_img
            c0 = 0
            family[0] += 1
            nc_all = _odd_chunks(units, 0)
            if WGRAD_TAP_SPLIT and units > 64 and nc_all <= 13:
                # all input channels at once (N = 8 * nc_all + 8, the spare column carries the bias gradient), taps 0..2 and 3..4:
                # (fields: ..., taps, n_cols, m_valid, n_valid, n0, s_m, s_n, s_t) + tap_shift as the last job field
                add(a_img, b_img, gp + 4 * w_off, gp + 4 * b_off, _lib.IMG_CHUNKS, 0, nc_all, 3, 8 * nc_all + 8, units, units, 0,
                    5 * units, 5, 1, -1)
                add(a_img, b_img, gp + 4 * (w_off + 3), None, _lib.IMG_CHUNKS, 0, nc_all, 2, 8 * nc_all + 8, units, units, 0,
                    5 * units, 5, 1, 2)
                continue
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
        family[0] += 1
        add(g, x, gp + 4 * w_off, gp + 4 * b_off, 1, 0, 1, 5, 16, units, cin0, 0, 5 * cin0, 5, 1)
        family[0] += 1
        add(y + (n_layer - 1) * layer_img, d, gp + 4 * lin_w_off, None, 1, 0, 1, 1, 16, units, fouts[st], 0, 1, units, 0)
    # The slabs of a layer all read the layer's whole gradient image (the A operand): they get the SAME group ranges and sit
    # next to each other in the list, so that they run at the same time on neighbouring SMs and the second reader of a window
    # finds it in L2 (measured before: 3.66 GB of DRAM reads per decoder step against 2.6 GB algorithmic).
    g_lo, g_hi = group_range if group_range is not None else (0, groups)
    if not (0 <= g_lo <= g_hi <= groups):
        raise ValueError("group_range %r outside [0, %d]" % (group_range, groups))
    n_g = g_hi - g_lo
    costs = [_job_cost_us(f[4], f[8], f[7]) * n_g for f, _ in protos]
    target = max(sum(costs) / (3.5 * (sm_count or n_sm())), 60.0)
    fam_cost = {}
    for (f, fam), c in zip(protos, costs):
        fam_cost[fam] = max(fam_cost.get(fam, 0.0), c)
    jobs = []
    for i, ((f, fam), c) in enumerate(zip(protos, costs)):
        n_split = splits if splits is not None else max(1, min(n_g, int(round(fam_cost[fam] / target))))
        for sp in range(n_split):
            g0, g1 = g_lo + n_g * sp // n_split, g_lo + n_g * (sp + 1) // n_split
            if g1 > g0:
                jobs.append(((-fam_cost[fam] * (g1 - g0) / max(n_g, 1), fam, sp, i), _lib.TaeWgradJob(*f[:15], g0, g1, f[15] if len(f) > 15 else 0)))
    jobs.sort(key=lambda t: t[0])
    return [j for _, j in jobs]


def pack_jobs(jobs, device):
    """(host array, count, device copy, device workspace): uploaded once, reused by every later launch."""
    arr = (_lib.TaeWgradJob * len(jobs))(*jobs)
    dev_copy = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(device)
    ws = torch.empty(512, dtype=torch.uint8, device=device)
    return arr, len(jobs), dev_copy, ws


def run_packed(packed, device):
    arr, n, dev_copy, ws = packed
    lib = _lib.load()
    _lib.check(lib.tae_wgrad_bf16(arr, n, _lib.ptr(dev_copy), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(device)))


def run_wgrad(jobs, device):
    packed = pack_jobs(jobs, device)
    run_packed(packed, device)
    return packed[3]


def backward_split(units: int, sm_count: int) -> int:
    """Work units that go into the FIRST of two backward launches, or 0 for one launch.  A unit (10 codewords of block length 100)
    occupies a CTA pair for the whole backward, so ``units`` units take ceil(units / pairs) waves; when the last wave is partly
    filled and the waves are few (batch 1000: 100 units on 74 pairs = one full wave + one 35 % full), the last wave becomes a
    launch of its own and the weight gradients of the earlier units run beside it on the SMs it leaves idle."""
    pairs = sm_count // 2
    if pairs < 1 or units <= pairs:
        return 0
    full, rest = divmod(units, pairs)
    if rest == 0 or full > 3 or rest > 0.7 * pairs:
        return 0
    return full * pairs


def _dec_offsets(a):
    F, units, n_layer, I = a.num_iter_ft, a.dec_num_unit, a.dec_num_layer, a.num_iteration
    offsets, fouts, off = [], [], 0
    for idx in range(I):
        for s_ in range(2):
            layers = []
            for j in range(n_layer):
                cin = (2 + F) if j == 0 else units
                layers.append((off, off + units * cin * 5))
                off += units * cin * 5 + units
            fout = 1 if (s_ == 1 and idx == I - 1) else F
            offsets.append((layers, off))
            fouts.append(fout)
            off += fout * units + fout
    return offsets, fouts


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
            buf = _buffers(dec, n_stacks, n_layer, groups, B, L, a.num_iter_ft, dev)
            tok, gen = buf.claim()
            perm, inv = dec.interleaver.device_index(dev)
            out = torch.empty((B, L, 1), dtype=torch.float32, device=dev)
            ws = dec._ws.get(256, dev)
            _lib.check(lib.tae_dec_forward_train_bf16(cfg, _lib.ptr(packed), _lib.ptr(received), _lib.ptr(perm), _lib.ptr(inv),
                                                      _lib.ptr(out), None, B, _lib.ptr(buf.stash_y), _lib.ptr(buf.stash_x),
                                                      _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)))
        ctx.dec, ctx.cfg, ctx.buf = dec, cfg, buf
        ctx.tok, ctx.gen = tok, gen
        ctx.shape = (B, L)
        ctx.need_input = received.requires_grad
        ctx.need_params = any(p.requires_grad for p in params)
        ctx.save_for_backward(out, flat)
        return out

    @staticmethod
    def backward(ctx, d_out):
        lib = _lib.load()
        dec, cfg, buf = ctx.dec, ctx.cfg, ctx.buf
        _check_stash(ctx)
        a = dec.args
        out, flat = ctx.saved_tensors
        B, L = ctx.shape
        dev = d_out.device
        F, units, n_layer, I = a.num_iter_ft, a.dec_num_unit, a.dec_num_layer, a.num_iteration
        n_stacks = 2 * I
        params = dec.ordered_parameters()
        with torch.cuda.device(dev):
            packed_bwd = dec._flat.derived.get("bf16_bwd")
            if packed_bwd is None:
                packed_bwd = torch.empty(lib.tae_dec_bwd_packed_bytes(cfg), dtype=torch.uint8, device=dev)
                _lib.check(lib.tae_dec_pack_bwd_bf16(cfg, _lib.ptr(flat), _lib.ptr(packed_bwd), _lib.stream_ptr(dev)))
                dec._flat.derived["bf16_bwd"] = packed_bwd
            perm, inv = dec.interleaver.device_index(dev)
            ws = dec._ws.get(256, dev)
            gflat, _ = _flat_grad(buf, flat, params)
            offsets, fouts = _dec_offsets(a)
            # out = sigmoid(deinterleave(o_last))  (decoders.py:267)  =>  d o_last = interleave(d_out * out * (1 - out))
            d_out = d_out.to(torch.float32).contiguous()
            d_o = torch.empty_like(out)
            _lib.check(lib.tae_dec_out_backward_f32(_lib.ptr(d_out), _lib.ptr(out), _lib.ptr(perm), _lib.ptr(d_o), B, L, _lib.stream_ptr(dev)))
            # all 2I stacks in one launch, schedule walked backwards (the glue between stacks runs inside the kernel) -- or in two
            # launches over disjoint work units when the last wave of units would leave most SMs idle (backward_split): the weight
            # gradients of the first launch's groups then run on a side stream beside the second launch
            n_units = lib.tae_train_units(L, B)
            cut = backward_split(n_units, n_sm(dev)) if (ctx.need_params and getattr(dec, "wgrad_overlap", WGRAD_OVERLAP)) else 0

            def run_backward(u0, u1):
                _lib.check(lib.tae_dec_backward_range_bf16(cfg, _lib.ptr(packed_bwd), _lib.ptr(d_o), _lib.ptr(perm), _lib.ptr(inv),
                                                           _lib.ptr(buf.stash_y), _lib.ptr(buf.stash_g), _lib.ptr(buf.stash_d),
                                                           _lib.ptr(buf.dxin), _lib.ptr(buf.dlin),
                                                           _lib.ptr(gflat) if ctx.need_params else None, B, u0, u1, _lib.ptr(ws), ws.numel(),
                                                           _lib.stream_ptr(dev)))
            if ctx.need_params and (buf.jobs is None or buf.jobs_cut != cut):
                mk = lambda rng: pack_jobs(wgrad_jobs(n_layer, units, 2 + F, fouts, buf.groups, buf.stash_y, buf.stash_x, buf.stash_g,
                                                      buf.stash_d, gflat, offsets, splits=getattr(dec, "wgrad_splits", None),
                                                      group_range=rng), dev)
                if cut:
                    buf.jobs, buf.jobs_tail = mk((0, 2 * cut)), mk((2 * cut, buf.groups))
                else:
                    buf.jobs, buf.jobs_tail = mk(None), None
                buf.jobs_cut = cut
            if cut:
                # (measured both ways: the second launch on this stream and the head's weight gradients on the side stream is a
                # little faster than the second launch on a high-priority side stream: 3.05 vs 3.13 ms per step at batch 1000)
                main = torch.cuda.current_stream(dev)
                if buf.side is None:
                    buf.side = torch.cuda.Stream(dev)
                run_backward(0, cut)
                first_done = torch.cuda.Event()
                first_done.record(main)
                run_backward(cut, n_units)                      # enqueued first: the latency-bound part
                buf.side.wait_event(first_done)                 # the groups of units [0, cut) are complete
                with torch.cuda.stream(buf.side):
                    run_packed(buf.jobs, dev)                   # ... their weight gradients run beside the second launch
                run_packed(buf.jobs_tail, dev)
                joined = torch.cuda.Event()
                joined.record(buf.side)
                main.wait_event(joined)
            else:
                run_backward(0, n_units)
                if ctx.need_params:
                    run_packed(buf.jobs, dev)
            d_rec = None
            if ctx.need_input:
                # stack inputs: even [r_sys, r_par1, prior] (decoders.py:230), odd [interleave(r_sys), r_par2, ...] (:240)
                d_rec = torch.empty((B, L, 3), dtype=torch.float32, device=dev)
                _lib.check(lib.tae_dec_input_grad_f32(_lib.ptr(buf.dxin), _lib.ptr(inv), _lib.ptr(d_rec), n_stacks, B, L, _lib.stream_ptr(dev)))
            grads = _grad_views(gflat, params) if ctx.need_params else [None] * len(params)
        ctx.tok.done = True
        return (None, d_rec, *grads)


def supported(args, which):
    """Static check (at module construction) whether the tensor-core training path covers this configuration."""
    if which == "dec":
        return (args.dec_kernel_size == 5 and 1 <= args.dec_num_unit <= 100 and args.dec_num_layer >= 2 and args.num_iter_ft <= 5
                and args.block_len <= 512)
    return args.enc_kernel_size == 5 and 1 <= args.enc_num_unit <= 100 and args.enc_num_layer >= 2 and args.block_len <= 512


def decoder_forward_train(dec, received):
    """Differentiable DEC_LargeCNN.forward on the tensor cores."""
    if dec.args.dec_kernel_size != 5:
        raise NotImplementedError("tensor-core training needs dec_kernel_size == 5")
    params = dec.ordered_parameters()
    return DecoderTrainFn.apply(dec, received.contiguous(), *params)


class EncoderTrainFn(torch.autograd.Function):
    """ENC_interCNN branches (reference encoders.py:362-373: three conv stacks, Linear(units, 1), ELU, concat) on the tensor
    cores: u (B, L, 1) -> un-normalised x_tx (B, L, 3) + its power statistics; power_constraint itself is shard.PowerNorm."""

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
            buf = _buffers(enc, 3, n_layer, groups, B, L, 1, dev)
            tok, gen = buf.claim()
            perm, inv = enc.interleaver.device_index(dev)
            x_tx = torch.empty((B, L, 3), dtype=torch.float32, device=dev)
            stats = torch.zeros(3, dtype=torch.float64, device=dev)
            ws = enc._ws.get(256, dev)
            _lib.check(lib.tae_enc_forward_train_bf16(cfg, _lib.ptr(packed), _lib.ptr(u), _lib.ptr(perm), _lib.ptr(inv), _lib.ptr(x_tx),
                                                      _lib.ptr(stats), B, _lib.ptr(buf.stash_y), _lib.ptr(buf.stash_x), _lib.ptr(ws),
                                                      ws.numel(), _lib.stream_ptr(dev)))
        ctx.enc, ctx.cfg, ctx.buf, ctx.shape = enc, cfg, buf, (B, L)
        ctx.tok, ctx.gen = tok, gen
        ctx.save_for_backward(x_tx, flat)
        ctx.mark_non_differentiable(stats)
        return x_tx, stats

    @staticmethod
    def backward(ctx, d_x, _d_stats=None):
        lib = _lib.load()
        enc, cfg, buf = ctx.enc, ctx.cfg, ctx.buf
        _check_stash(ctx)
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
            params = enc.ordered_parameters()
            gflat, _ = _flat_grad(buf, flat, params)
            # x_tx = ELU(Linear(h))  =>  d lin = d x_tx * ELU'   (ELU' = x_tx + 1 where x_tx < 0); one (B, L, 1) slab per branch
            d_x = d_x.to(torch.float32).contiguous()
            d_lin = torch.empty((3, B, L, 1), dtype=torch.float32, device=dev)
            _lib.check(lib.tae_enc_out_backward_f32(_lib.ptr(d_x), _lib.ptr(x_tx), _lib.ptr(d_lin), B, L, _lib.stream_ptr(dev)))
            offsets, off = [], 0
            for br in range(3):
                layers = []
                for j in range(n_layer):
                    cin = 1 if j == 0 else units
                    layers.append((off, off + units * cin * 5))
                    off += units * cin * 5 + units
                offsets.append((layers, off))
                off += units + 1
            _lib.check(lib.tae_enc_backward_bf16(cfg, _lib.ptr(packed_bwd), _lib.ptr(d_lin), _lib.ptr(buf.stash_y), _lib.ptr(buf.stash_g),
                                                 _lib.ptr(buf.stash_d), _lib.ptr(buf.dxin), _lib.ptr(gflat), B, _lib.ptr(ws), ws.numel(),
                                                 _lib.stream_ptr(dev)))
            if buf.jobs is None:
                jobs = wgrad_jobs(n_layer, units, 1, [1, 1, 1], buf.groups, buf.stash_y, buf.stash_x, buf.stash_g, buf.stash_d, gflat,
                                  offsets, splits=getattr(enc, "wgrad_splits", None))
                buf.jobs = pack_jobs(jobs, dev)
            run_packed(buf.jobs, dev)
            grads = _grad_views(gflat, params)
        ctx.tok.done = True
        return (None, None, *grads)


def encoder_branches_train(enc, u):
    """Differentiable un-normalised ENC_interCNN output (B, L, 3) on the tensor cores, and the (sum, sum of squares, count) of its
    elements as 3 device doubles (what power_constraint needs: encoders.py:107-116)."""
    if enc.args.enc_kernel_size != 5:
        raise NotImplementedError("tensor-core training needs enc_kernel_size == 5")
    return EncoderTrainFn.apply(enc, u.contiguous(), *enc.ordered_parameters())
