"""DEC_LargeCNN with the reference's nn.Module surface (reference decoders.py:157-269).

forward: ``received (B, L, 3)`` -> ``(B, L, 1)`` posteriors.  The whole turbo schedule (2*num_iteration conv
stacks, Linear projections, extrinsic subtractions, interleave / de-interleave, sigmoid) is one call into
libturboae_b200.so: ``precision='bf16'`` runs the fused tcgen05 kernel, ``precision='f16x3'`` the split-operand
tcgen05 kernel (elementwise parity), ``precision='fp32'`` the CUDA-core elementwise-parity path.  No CPU fallback and no silent switch between the two."""
from __future__ import annotations

import os

import torch

from . import _lib
from ._flat import FlatCache, OrderedParameters, ParallelShim, Workspace, unwrap
from ._flat import _EPOCH as _flat_epoch_cell


def _flat_epoch():
    return _flat_epoch_cell[0]
from .cnn_utils import DenseSameShapeConv1d, SameShapeConv1d
from .interleavers import DeInterleaver, Interleaver


class DEC_LargeCNN(OrderedParameters, torch.nn.Module):
    def __init__(self, args, p_array):
        super().__init__()
        self.args = args
        use_cuda = not args.no_cuda and torch.cuda.is_available()
        self.this_device = torch.device("cuda" if use_cuda else "cpu")
        # decoders.py:173 keys the layer type on args.encoder: the dense variant runs layer by layer on the fp32 kernels (the
        # fused tensor-core schedule and the flat-parameter C ABI cover the SameShapeConv1d layout only)
        self.dense = args.encoder != "TurboAE_rate3_cnn"
        CNNLayer = DenseSameShapeConv1d if self.dense else SameShapeConv1d
        self.interleaver = Interleaver(args, p_array)
        self.deinterleaver = DeInterleaver(args, p_array)
        self.dec1_cnns = torch.nn.ModuleList()
        self.dec2_cnns = torch.nn.ModuleList()
        self.dec1_outputs = torch.nn.ModuleList()
        self.dec2_outputs = torch.nn.ModuleList()
        for idx in range(args.num_iteration):
            for lst in (self.dec1_cnns, self.dec2_cnns):
                lst.append(CNNLayer(num_layer=args.dec_num_layer, in_channels=2 + args.num_iter_ft,
                                    out_channels=args.dec_num_unit, kernel_size=args.dec_kernel_size))
            self.dec1_outputs.append(torch.nn.Linear(args.dec_num_unit, args.num_iter_ft))
            self.dec2_outputs.append(torch.nn.Linear(args.dec_num_unit,
                                                     1 if idx == args.num_iteration - 1 else args.num_iter_ft))
        self._flat = FlatCache()
        self._watch_ordered()
        self._ws = Workspace()
        self._ws_host = Workspace()
        #: 'bf16' (fused tcgen05 kernel: BER parity), 'f16x3' (split-operand tcgen05 kernel: elementwise parity <= 1e-4 at
        #: ~13x the fp32 rate) or 'fp32' (CUDA-core parity path)
        #: 'auto' (default) = 'bf16' wherever the fused kernel covers the configuration (kernel size 5, <= 100 units, block length
        #: <= 512), else the fp32 kernels -- said once in a warning, never silent
        self.precision = getattr(args, "tae_precision", None) or os.environ.get("TURBOAE_B200_PRECISION", "auto")
        #: training (autograd) path: 'fp32' = CUDA-core kernels layer by layer (gradients within 2e-3 of the reference's),
        #: 'bf16' = tensor cores (train_tc.py: fused forward with stash, fused backward per stack, weight-gradient GEMMs)
        #: default: 'bf16' whenever the tensor path covers the configuration (decided here, once, from args)
        from . import train_tc
        self.train_precision = (getattr(args, "tae_train_precision", None) or os.environ.get("TURBOAE_B200_TRAIN_PRECISION")
                                or ("bf16" if (train_tc.supported(args, "dec") and not self.dense) else "fp32"))

    def set_parallel(self):
        self._drop_ordered()
        for lst in (self.dec1_cnns, self.dec2_cnns, self.dec1_outputs, self.dec2_outputs):
            for idx in range(len(lst)):
                if not isinstance(lst[idx], ParallelShim):
                    lst[idx] = ParallelShim(lst[idx])

    def set_interleaver(self, p_array):
        self.interleaver.set_parray(p_array)
        self.deinterleaver.set_parray(p_array)

    # -- canonical flat order of include/turboae_b200.h ------------------------------------------------
    def _walk_ordered_parameters(self):
        out = []
        for idx in range(self.args.num_iteration):
            for cnns, outs in ((self.dec1_cnns, self.dec1_outputs), (self.dec2_cnns, self.dec2_outputs)):
                for conv in unwrap(cnns[idx]).cnns:
                    out += [conv.weight, conv.bias]
                lin = unwrap(outs[idx])
                out += [lin.weight, lin.bias]
        return out

    def config(self, block_len):
        a = self.args
        return _lib.TaeDecConfig(block_len, a.num_iteration, a.num_iter_ft, a.dec_num_layer, a.dec_num_unit,
                                 a.dec_kernel_size, 1 if a.extrinsic else 0)

    def decode(self, received, precision=None, trace=None):
        """received: contiguous float32 CUDA (B, L, 3) -> (B, L, 1).  `trace`: optional (2I, B, L, F) tensor that
        receives every dec{1,2}_outputs Linear output (pre-subtraction), for parity debugging."""
        _lib.require_cuda(received, "DEC_LargeCNN input")
        B, L, three = received.shape
        if three != 3:
            raise _lib.TaeError("DEC_LargeCNN expects (B, L, 3), got %s" % (tuple(received.shape),))
        dev = received.device
        out = torch.empty((B, L, 1), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            lib, cfg, flat, packed, prec = self._prepare(L, dev, precision)
            if flat.numel() != lib.tae_dec_param_count(cfg):
                raise _lib.TaeError("parameter count mismatch: module %d vs library %d"
                                    % (flat.numel(), lib.tae_dec_param_count(cfg)))
            perm, inv = self.interleaver.device_index(dev)
            ws_bytes = lib.tae_dec_workspace_bytes(cfg, B, prec)
            ws = self._ws.get(ws_bytes, dev)
            _lib.check(lib.tae_dec_forward(cfg, _lib.ptr(flat), _lib.ptr(packed), _lib.ptr(received), _lib.ptr(perm),
                                           _lib.ptr(inv), _lib.ptr(out), _lib.ptr(trace), B, prec, _lib.ptr(ws),
                                           ws.numel(), _lib.stream_ptr(dev)))
        return out

    def resolved_precision(self, block_len, precision=None):
        """The inference path a forward at this block length takes ('auto' resolved; see ``precision``)."""
        precision = precision or self.precision
        if precision != "auto":
            if precision not in _lib.PRECISIONS:
                raise _lib.TaeError("precision must be 'auto', 'bf16', 'f16x3' or 'fp32', got %r" % (precision,))
            return precision
        lib = _lib.load()
        if lib.tae_dec_packed_bytes(self.config(block_len)):
            return "bf16"
        if not getattr(self, "_warned_fp32", False):          # said once: the choice is by shape, never silent
            import warnings
            warnings.warn("turboae_b200.DEC_LargeCNN: the fused tensor-core kernel does not cover this configuration (%s); "
                          "using the fp32 CUDA-core kernels (elementwise parity, ~45x slower)" % lib.tae_last_error().decode())
            self._warned_fp32 = True
        return "fp32"

    def _prepare(self, L, dev, precision):
        lib = _lib.load()
        precision = self.resolved_precision(L, precision)
        prec = _lib.PRECISIONS[precision]
        cfg = self.config(L)
        flat = self._flat.get(self.ordered_parameters())
        if flat.device != dev:
            raise _lib.TaeError("decoder parameters are on %s but the device buffers are on %s" % (flat.device, dev))
        packed = None
        if prec == _lib.PRECISION_BF16:
            packed = self._flat.derived.get("bf16")
            if packed is None:
                nbytes = lib.tae_dec_packed_bytes(cfg)
                if nbytes == 0:
                    raise _lib.TaeError("bf16 tensor path unavailable for this configuration (%s); set precision='fp32'"
                                        % lib.tae_last_error().decode())
                packed = torch.empty(nbytes, dtype=torch.uint8, device=dev)
                _lib.check(lib.tae_dec_pack_bf16(cfg, _lib.ptr(flat), _lib.ptr(packed), _lib.stream_ptr(dev)))
                self._flat.derived["bf16"] = packed
        elif prec == _lib.PRECISION_F16X3:
            packed = self._flat.derived.get("f16x3")
            if packed is None:
                nbytes = lib.tae_dec_packed_bytes_x3(cfg)
                if nbytes == 0:
                    raise _lib.TaeError("f16x3 tensor path unavailable for this configuration (%s); set precision='fp32'"
                                        % lib.tae_last_error().decode())
                packed = torch.empty(nbytes, dtype=torch.uint8, device=dev)
                _lib.check(lib.tae_dec_pack_f16x3(cfg, _lib.ptr(flat), _lib.ptr(packed), _lib.stream_ptr(dev)))
                self._flat.derived["f16x3"] = packed
        return lib, cfg, flat, packed, prec

    def decode_host(self, received_host, out_host=None, precision=None):
        """Host-buffer decode: ``received_host`` (B, L, 3) float32 on the CPU (pin it for asynchronous copies) ->
        ``out_host`` (B, L, 1) on the CPU.  H2D copy, fused decode and D2H copy of successive chunks of the batch overlap
        inside ``tae_dec_forward_host``; the call returns after enqueueing, synchronise the current stream (or call
        ``torch.cuda.synchronize()``) before reading ``out_host``."""
        if received_host.is_cuda or received_host.dtype != torch.float32 or not received_host.is_contiguous():
            raise _lib.TaeError("decode_host expects a contiguous float32 CPU tensor")
        if self.this_device.type != "cuda":
            raise _lib.TaeError("no CUDA device: turboae_b200 has no CPU fallback")
        B, L, three = received_host.shape
        if three != 3:
            raise _lib.TaeError("DEC_LargeCNN expects (B, L, 3), got %s" % (tuple(received_host.shape),))
        dev = next(self.parameters()).device
        if out_host is None:
            out_host = torch.empty((B, L, 1), dtype=torch.float32).pin_memory()
        if out_host.is_cuda or out_host.dtype != torch.float32 or out_host.numel() != B * L or not out_host.is_contiguous():
            raise _lib.TaeError("out_host must be a contiguous float32 CPU tensor with B*L elements")
        with torch.cuda.device(dev):
            lib, cfg, flat, packed, prec = self._prepare(L, dev, precision)
            perm, inv = self.interleaver.device_index(dev)
            ws_bytes = lib.tae_dec_host_workspace_bytes(cfg, B, prec)
            ws = self._ws_host.get(ws_bytes, dev)
            _lib.check(lib.tae_dec_forward_host(cfg, _lib.ptr(flat), _lib.ptr(packed), _lib.ptr(received_host), _lib.ptr(perm),
                                                _lib.ptr(inv), _lib.ptr(out_host), B, prec, _lib.ptr(ws), ws.numel(),
                                                _lib.stream_ptr(dev)))
        return out_host

    def _forward_train(self, received):
        """Autograd path (reference trainer.py:64-76 backpropagates through decoders.py:219-269): the turbo schedule is
        spelled out with this package's differentiable pieces -- SameShapeConv1d (fp32 conv + ELU kernels forward, weight- and
        data-gradient kernels backward: ~98 % of the FLOPs), Interleaver / DeInterleaver (gather kernel both ways) -- and
        torch glue for the concat, the 100->5 Linear and the subtractions."""
        a = self.args
        B, L, _ = received.shape
        r_sys, r_par1, r_par2 = received[:, :, 0:1], received[:, :, 1:2], received[:, :, 2:3]
        r_sys_int = self.interleaver(r_sys)
        prior = torch.zeros((B, L, a.num_iter_ft), dtype=torch.float32, device=received.device)
        x_plr = None
        for idx in range(a.num_iteration):
            last = idx == a.num_iteration - 1
            x_plr = self.dec1_outputs[idx](self.dec1_cnns[idx](torch.cat([r_sys, r_par1, prior], dim=2)))
            if a.extrinsic:
                x_plr = x_plr - prior
            x_plr_int = self.interleaver(x_plr)
            x_plr = self.dec2_outputs[idx](self.dec2_cnns[idx](torch.cat([r_sys_int, r_par2, x_plr_int], dim=2)))
            if not last:
                if a.extrinsic:
                    x_plr = x_plr - x_plr_int
                prior = self.deinterleaver(x_plr)
        return torch.sigmoid(self.deinterleaver(x_plr))

    def forward(self, received):
        a = self.args
        if getattr(a, "is_variable_block_len", False) and a.is_interleave != 0:
            # reference decoders.py:208-215: the interleaver is re-drawn for the length of this batch (same draw from
            # numpy's global generator as the reference)
            import numpy as np
            seed = np.random.randint(0, a.is_interleave)
            self.set_interleaver(np.random.mtrand.RandomState(seed).permutation(np.arange(received.shape[1])))
        if self.this_device.type != "cuda":
            raise _lib.TaeError("no CUDA device: turboae_b200 has no CPU fallback")
        if self.dense:
            # DenseSameShapeConv1d stacks: the schedule spelled out with the module pieces (fp32 kernels), with or without autograd
            return self._forward_train(received.to(device=self.this_device, dtype=torch.float32).contiguous())
        if torch.is_grad_enabled() and (received.requires_grad or any(p.requires_grad for p in self.parameters())):
            x = received.to(device=self.this_device, dtype=torch.float32)
            if self.train_precision == "bf16":
                from . import train_tc
                return train_tc.decoder_forward_train(self, x)
            if self.train_precision != "fp32":
                raise _lib.TaeError("train_precision must be 'bf16' or 'fp32', got %r" % (self.train_precision,))
            return self._forward_train(x)
        # reference decoders.py:219: received.type(torch.FloatTensor).to(self.this_device) -- here without the
        # device->host->device round trip when the tensor is already resident.
        x = received.to(device=self.this_device, dtype=torch.float32).contiguous()
        return self.decode(x)


class _GruDirectionFn(torch.autograd.Function):
    """One direction of one GRU layer under autograd (training of DEC_LargeRNN, reference trainer.py:74 through
    decoders.py:86-149): forward = input projection (K = 1 case of the conv kernel) + ``tae_gru_direction_f32``; backward =
    ``tae_gru_direction_bwd_f32`` (the sequential part: back-propagation through time with the gates recomputed) followed by
    the reductions over (B, L), which are plain GEMMs (torch.matmul -> cuBLAS): dW_ih = dgi^T x, dx = dgi W_ih,
    dW_hh = dgh^T h_prev, and the bias sums."""

    @staticmethod
    def forward(ctx, x, w_ih, w_hh, b_ih, b_hh, reverse):
        lib = _lib.load()
        B, L, _ = x.shape
        H = w_hh.shape[1]
        x = x.contiguous()
        w_ih_c, w_hh_c, b_ih_c, b_hh_c = (t.detach().contiguous() for t in (w_ih, w_hh, b_ih, b_hh))
        with torch.cuda.device(x.device):
            xproj = DEC_LargeRNN._pointwise(x, w_ih_c, b_ih_c)
            out = torch.empty((B, L, H), dtype=torch.float32, device=x.device)
            _lib.check(lib.tae_gru_direction_f32(_lib.ptr(xproj), _lib.ptr(w_hh_c), _lib.ptr(b_hh_c), _lib.ptr(out), B, L, H, H, 0,
                                                 1 if reverse else 0, _lib.stream_ptr(x.device)))
        ctx.reverse = bool(reverse)
        ctx.save_for_backward(x, xproj, out, w_ih_c, w_hh_c, b_hh_c)
        return out

    @staticmethod
    def backward(ctx, d_out):
        lib = _lib.load()
        x, xproj, out, w_ih, w_hh, b_hh = ctx.saved_tensors
        B, L, H = out.shape
        d_out = d_out.contiguous()
        with torch.cuda.device(x.device):
            dgi = torch.empty((B, L, 3 * H), dtype=torch.float32, device=x.device)
            dghn = torch.empty((B, L, H), dtype=torch.float32, device=x.device)
            _lib.check(lib.tae_gru_direction_bwd_f32(_lib.ptr(xproj), _lib.ptr(w_hh), _lib.ptr(b_hh), _lib.ptr(out), _lib.ptr(d_out),
                                                     _lib.ptr(dgi), _lib.ptr(dghn), B, L, H, H, 0, 1 if ctx.reverse else 0,
                                                     _lib.stream_ptr(x.device)))
        g2 = dgi.reshape(B * L, 3 * H)
        h_prev = torch.zeros_like(out)                    # h_{t-1} in the direction of the recurrence
        if ctx.reverse:
            h_prev[:, :-1] = out[:, 1:]
        else:
            h_prev[:, 1:] = out[:, :-1]
        dgh = torch.cat([dgi[:, :, :2 * H], dghn], dim=2).reshape(B * L, 3 * H)
        dx = (g2 @ w_ih).reshape(B, L, -1) if ctx.needs_input_grad[0] else None
        dw_ih = g2.t() @ x.reshape(B * L, -1)
        dw_hh = dgh.t() @ h_prev.reshape(B * L, H)
        return dx, dw_ih, dw_hh, g2.sum(0), dgh.sum(0), None


class DEC_LargeRNN(torch.nn.Module):
    """DeepTurbo decoder with the reference's nn.Module surface (reference decoders.py:16-149): 2*num_iteration stacks of a
    2-layer bidirectional GRU(2+F -> H) + Linear(2H -> F).  The ``torch.nn.GRU`` children only hold the parameters (same names
    and shapes as the reference, so its checkpoints load); forward runs ``tae_conv1d_elu_f32`` (K = 1: all input projections
    of a layer-direction as one pointwise GEMM), ``tae_gru_direction_f32`` (the recurrence, W_hh resident in shared memory)
    and the interleaver gather; ``precision='bf16'`` (default) runs the recurrence on the tensor cores.  Under autograd
    (training, reference trainer.py:74) every GRU direction goes through ``_GruDirectionFn`` (fp32 recurrence forward,
    ``tae_gru_direction_bwd_f32`` backward)."""

    def __init__(self, args, p_array):
        super().__init__()
        self.args = args
        use_cuda = not args.no_cuda and torch.cuda.is_available()
        self.this_device = torch.device("cuda" if use_cuda else "cpu")
        if getattr(args, "dec_rnn", "gru") != "gru":
            raise NotImplementedError("turboae_b200.DEC_LargeRNN implements dec_rnn='gru' only (got %r)" % (args.dec_rnn,))
        if getattr(args, "dec_act", "linear") != "linear":
            raise NotImplementedError("dec_act=%r: only 'linear' is built" % (args.dec_act,))
        self.interleaver = Interleaver(args, p_array)
        self.deinterleaver = DeInterleaver(args, p_array)
        self.dec1_rnns, self.dec2_rnns = torch.nn.ModuleList(), torch.nn.ModuleList()
        self.dec1_outputs, self.dec2_outputs = torch.nn.ModuleList(), torch.nn.ModuleList()
        for idx in range(args.num_iteration):
            for lst in (self.dec1_rnns, self.dec2_rnns):
                lst.append(torch.nn.GRU(2 + args.num_iter_ft, args.dec_num_unit, num_layers=2, bias=True, batch_first=True,
                                        dropout=getattr(args, "dropout", 0.0), bidirectional=True))
            self.dec1_outputs.append(torch.nn.Linear(2 * args.dec_num_unit, args.num_iter_ft))
            self.dec2_outputs.append(torch.nn.Linear(2 * args.dec_num_unit, 1 if idx == args.num_iteration - 1 else args.num_iter_ft))
        H = args.dec_num_unit
        tc_ok = H % 4 == 0 and 4 <= H <= 100 and 2 + args.num_iter_ft <= 200
        #: 'bf16' = tensor-core recurrence (tae_gru_direction_bf16: bf16 operands, fp32 accumulation and state), 'fp32' = CUDA-core
        #: recurrence (elementwise parity with torch.nn.GRU); the default is decided here, once, from the configuration
        self.precision = (getattr(args, "tae_rnn_precision", None) or os.environ.get("TURBOAE_B200_RNN_PRECISION")
                          or ("bf16" if tc_ok else "fp32"))
        self._packed = {}
        self._ws = Workspace()

    def set_parallel(self):
        for lst in (self.dec1_rnns, self.dec2_rnns, self.dec1_outputs, self.dec2_outputs):
            for idx in range(len(lst)):
                if not isinstance(lst[idx], ParallelShim):
                    lst[idx] = ParallelShim(lst[idx])

    def set_interleaver(self, p_array):
        self.interleaver.set_parray(p_array)
        self.deinterleaver.set_parray(p_array)

    @staticmethod
    def _pointwise(x, w, b):
        """(B, L, Cin) @ w(Cout, Cin)^T + b through the K = 1 case of the conv kernel."""
        lib = _lib.load()
        B, L, cin = x.shape
        cout = w.shape[0]
        ws_bytes = lib.tae_conv1d_workspace_bytes(cin, cout, 1)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
        out = torch.empty((B, L, cout), dtype=torch.float32, device=x.device)
        _lib.check(lib.tae_conv1d_elu_f32(_lib.ptr(x), _lib.ptr(out), _lib.ptr(w), _lib.ptr(b), B, L, cin, cout, 1, 0, _lib.ptr(ws),
                                          ws_bytes, _lib.stream_ptr(x.device)))
        return out

    def _stack_tc(self, gru, lin, x):
        """One decoder stack on the tensor cores: 2-layer bidirectional GRU + its Linear, (B, L, 2+F) fp32 -> (B, L, Fout) fp32.
        Activations stay in time-major bf16 tiles between the launches (include/turboae_b200.h)."""
        lib = _lib.load()
        gru, lin = unwrap(gru), unwrap(lin)
        H = gru.hidden_size
        B, L, cin = x.shape
        dev = x.device
        R = lib.tae_gru_rows_per_block(B)
        ws = self._ws.get(256, dev)
        stream = _lib.stream_ptr(dev)
        tiles = torch.empty(lib.tae_gru_tile_bytes(B, L, (cin + 7) // 8, R), dtype=torch.uint8, device=dev)
        _lib.check(lib.tae_gru_tiles_from_f32(_lib.ptr(x.contiguous()), _lib.ptr(tiles), B, L, cin, R, stream))
        in_ch, grp = cin, cin
        n_out = 2 * ((H + 7) // 8)
        for layer in range(gru.num_layers):
            out = torch.empty(lib.tae_gru_tile_bytes(B, L, n_out, R), dtype=torch.uint8, device=dev)
            for d, suffix in enumerate(("", "_reverse")):
                k = "l%d%s" % (layer, suffix)
                ps = [getattr(gru, n + k) for n in ("weight_ih_", "weight_hh_", "bias_ih_", "bias_hh_")]
                key = (id(gru), k)
                ver = (_flat_epoch(),) + tuple((p.data_ptr(), p._version) for p in ps)     # epoch: writes through .data (_flat.py)
                ent = self._packed.get(key)
                if ent is None or ent[0] != ver:
                    nbytes = lib.tae_gru_packed_bytes(H, in_ch, grp)
                    if nbytes == 0:
                        raise _lib.TaeError("bf16 GRU path unavailable (%s); set precision='fp32'" % lib.tae_last_error().decode())
                    packed = torch.empty(nbytes, dtype=torch.uint8, device=dev)
                    w = [p.detach().to(torch.float32).contiguous() for p in ps]
                    _lib.check(lib.tae_gru_pack_bf16(_lib.ptr(w[0]), _lib.ptr(w[1]), _lib.ptr(w[2]), _lib.ptr(w[3]), _lib.ptr(packed), H, in_ch,
                                                     grp, stream))
                    ent = (ver, packed, w)
                    self._packed[key] = ent
                _lib.check(lib.tae_gru_direction_bf16(_lib.ptr(ent[1]), _lib.ptr(tiles), _lib.ptr(out), B, L, H, in_ch, grp, R, n_out,
                                                      d * (n_out // 2), d, _lib.ptr(ws), ws.numel(), stream))
            tiles, in_ch, grp = out, 2 * H, H
        F = lin.out_features
        y = torch.empty((B, L, F), dtype=torch.float32, device=dev)
        _lib.check(lib.tae_gru_linear_f32(_lib.ptr(tiles), _lib.ptr(lin.weight.detach().contiguous()), _lib.ptr(lin.bias.detach().contiguous()),
                                          _lib.ptr(y), B, L, 2 * H, H, F, R, stream))
        return y

    def _stack(self, gru, lin, x):
        if self.precision == "bf16":
            return self._stack_tc(gru, lin, x)
        if self.precision != "fp32":
            raise _lib.TaeError("precision must be 'bf16' or 'fp32', got %r" % (self.precision,))
        lin = unwrap(lin)
        return self._pointwise(self._gru_stack(gru, x), lin.weight.detach().contiguous(), lin.bias.detach().contiguous())

    def _gru_stack(self, gru, x):
        lib = _lib.load()
        gru = unwrap(gru)
        H = gru.hidden_size
        B, L, _ = x.shape
        h = x.contiguous()
        for layer in range(gru.num_layers):
            out = torch.empty((B, L, 2 * H), dtype=torch.float32, device=x.device)
            for d, suffix in enumerate(("", "_reverse")):
                k = "l%d%s" % (layer, suffix)
                w_ih, w_hh = getattr(gru, "weight_ih_" + k).detach().contiguous(), getattr(gru, "weight_hh_" + k).detach().contiguous()
                b_ih, b_hh = getattr(gru, "bias_ih_" + k).detach().contiguous(), getattr(gru, "bias_hh_" + k).detach().contiguous()
                xproj = self._pointwise(h, w_ih, b_ih)
                _lib.check(lib.tae_gru_direction_f32(_lib.ptr(xproj), _lib.ptr(w_hh), _lib.ptr(b_hh), _lib.ptr(out), B, L, H, 2 * H,
                                                     d * H, d, _lib.stream_ptr(x.device)))
            h = out
        return h

    def _stack_train(self, gru, lin, x):
        """One decoder stack under autograd: 2-layer bidirectional GRU (fp32 recurrence kernels, _GruDirectionFn) + Linear, with
        the reference's dropout placement (between the GRU layers: torch.nn.GRU(dropout=...), decoders.py:43-52; on the Linear
        output: decoders.py:103)."""
        gru, lin = unwrap(gru), unwrap(lin)
        p_drop = float(getattr(self.args, "dropout", 0.0))
        h = x
        for layer in range(gru.num_layers):
            outs = []
            for d, suffix in enumerate(("", "_reverse")):
                k = "l%d%s" % (layer, suffix)
                outs.append(_GruDirectionFn.apply(h, getattr(gru, "weight_ih_" + k), getattr(gru, "weight_hh_" + k),
                                                  getattr(gru, "bias_ih_" + k), getattr(gru, "bias_hh_" + k), d == 1))
            h = torch.cat(outs, dim=2)
            if layer < gru.num_layers - 1 and p_drop > 0.0:
                h = torch.nn.functional.dropout(h, p_drop, self.training)
        y = torch.nn.functional.linear(h, lin.weight, lin.bias)
        return torch.nn.functional.dropout(y, p_drop, self.training) if p_drop > 0.0 else y

    def _forward_train(self, received):
        """Autograd path (reference trainer.py:64-76 through decoders.py:86-149): the same turbo schedule as forward(), every
        stack through _stack_train; interleave / de-interleave are this package's differentiable gathers."""
        a = self.args
        B, L, _ = received.shape
        r_sys, r_par1, r_par2 = received[:, :, 0:1], received[:, :, 1:2], received[:, :, 2:3]
        r_sys_int = self.interleaver(r_sys.contiguous())
        prior = torch.zeros((B, L, a.num_iter_ft), dtype=torch.float32, device=received.device)
        x_plr = None
        for idx in range(a.num_iteration):
            last = idx == a.num_iteration - 1
            x_plr = self._stack_train(self.dec1_rnns[idx], self.dec1_outputs[idx], torch.cat([r_sys, r_par1, prior], dim=2))
            if a.extrinsic:
                x_plr = x_plr - prior
            x_plr_int = self.interleaver(x_plr.contiguous())
            x_plr = self._stack_train(self.dec2_rnns[idx], self.dec2_outputs[idx], torch.cat([r_sys_int, r_par2, x_plr_int], dim=2))
            if not last:
                if a.extrinsic:
                    x_plr = x_plr - x_plr_int
                prior = self.deinterleaver(x_plr.contiguous())
        return torch.sigmoid(self.deinterleaver(x_plr.contiguous()))

    def forward(self, received):
        if self.this_device.type != "cuda":
            raise _lib.TaeError("no CUDA device: turboae_b200 has no CPU fallback")
        a = self.args
        received = received.to(device=self.this_device, dtype=torch.float32).contiguous()
        if torch.is_grad_enabled() and (received.requires_grad or any(p.requires_grad for p in self.parameters())):
            return self._forward_train(received)
        B, L, _ = received.shape
        with torch.cuda.device(received.device):
            r_sys, r_par1, r_par2 = received[:, :, 0:1], received[:, :, 1:2], received[:, :, 2:3]
            r_sys_int = self.interleaver(r_sys.contiguous())
            prior = torch.zeros((B, L, a.num_iter_ft), dtype=torch.float32, device=received.device)
            x_plr = None
            for idx in range(a.num_iteration):
                last = idx == a.num_iteration - 1
                x_plr = self._stack(self.dec1_rnns[idx], self.dec1_outputs[idx], torch.cat([r_sys, r_par1, prior], dim=2))
                if a.extrinsic:
                    x_plr = x_plr - prior
                x_plr_int = self.interleaver(x_plr)
                x_plr = self._stack(self.dec2_rnns[idx], self.dec2_outputs[idx], torch.cat([r_sys_int, r_par2, x_plr_int], dim=2))
                if not last:
                    if a.extrinsic:
                        x_plr = x_plr - x_plr_int
                    prior = self.deinterleaver(x_plr)
            return torch.sigmoid(self.deinterleaver(x_plr))
