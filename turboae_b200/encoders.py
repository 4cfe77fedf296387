"""ENC_interCNN with the reference's nn.Module surface (reference encoders.py:63-125, 306-377).

forward: ``u (B, L, 1) in {0,1}`` -> ``codes (B, L, 3)``: three conv branches (the third on interleaved bits),
Linear(100->1) + ELU each, concat, then the batch-global power normalisation -- all in libturboae_b200.so
(``tae_enc_forward`` + ``tae_power_norm_f32``)."""
from __future__ import annotations

import os

import torch

from . import _lib, shard
from ._flat import FlatCache, OrderedParameters, ParallelShim, Workspace, unwrap
from .cnn_utils import DenseSameShapeConv1d, SameShapeConv1d
from .interleavers import Interleaver


class STEQuantize(torch.autograd.Function):
    """reference encoders.py:20-57: quantised forward, clipped straight-through backward (torch glue: elementwise on (B,L,3)).
    The experimental 'group_norm_noisy' gradient-noise branch (encoders.py:48-55) is not supported."""

    @staticmethod
    def forward(ctx, inputs, args):
        ctx.save_for_backward(inputs)
        ctx.args = args
        lim = args.enc_value_limit
        x = torch.clamp(inputs, -lim, lim)
        if args.enc_quantize_level == 2:
            return torch.sign(x)
        q = args.enc_quantize_level
        return torch.round((x + lim) * ((q - 1.0) / (2.0 * lim))) * (2.0 * lim) / (q - 1.0) - lim

    @staticmethod
    def backward(ctx, grad_output):
        a = ctx.args
        g = grad_output.clone()
        if a.enc_clipping in ("inputs", "both"):
            inp, = ctx.saved_tensors
            g[inp > a.enc_value_limit] = 0
            g[inp < -a.enc_value_limit] = 0
        if a.enc_clipping in ("gradient", "both"):
            g = torch.clamp(g, -a.enc_grad_limit, a.enc_grad_limit)
        return g, None


class ENCBase(torch.nn.Module):
    """reference encoders.py:63-125."""

    def __init__(self, args):
        super().__init__()
        use_cuda = not args.no_cuda and torch.cuda.is_available()
        self.this_device = torch.device("cuda" if use_cuda else "cpu")
        self.args = args
        self.reset_precomp()

    def set_parallel(self):
        pass

    def set_precomp(self, mean_scalar, std_scalar):
        self.mean_scalar = mean_scalar.to(self.this_device)
        self.std_scalar = std_scalar.to(self.this_device)

    def reset_precomp(self):
        self.mean_scalar = torch.zeros(1).type(torch.FloatTensor).to(self.this_device)
        self.std_scalar = torch.ones(1).type(torch.FloatTensor).to(self.this_device)
        self.num_test_block = 0.0


class ENC_interCNN(OrderedParameters, ENCBase):
    """reference encoders.py:306-377."""

    def __init__(self, args, p_array):
        super().__init__(args)
        self.args = args
        if args.code_rate_k != 1:
            raise NotImplementedError("code_rate_k must be 1 (got %r)" % (args.code_rate_k,))
        # encoders.py:313-330 keys the layer type on args.encoder: the dense variant (-encoder TurboAE_rate3_cnn_dense) runs layer
        # by layer on the fp32 kernels (the fused tensor-core kernels and the flat-parameter C ABI cover SameShapeConv1d only)
        self.dense = args.encoder != "TurboAE_rate3_cnn"
        CNNLayer = DenseSameShapeConv1d if self.dense else SameShapeConv1d
        mk = lambda: CNNLayer(num_layer=args.enc_num_layer, in_channels=args.code_rate_k,
                              out_channels=args.enc_num_unit, kernel_size=args.enc_kernel_size)
        self.enc_cnn_1, self.enc_cnn_2, self.enc_cnn_3 = mk(), mk(), mk()
        self.enc_linear_1 = torch.nn.Linear(args.enc_num_unit, 1)
        self.enc_linear_2 = torch.nn.Linear(args.enc_num_unit, 1)
        self.enc_linear_3 = torch.nn.Linear(args.enc_num_unit, 1)
        self.interleaver = Interleaver(args, p_array)
        self._flat = FlatCache()
        self._watch_ordered()
        self._ws = Workspace()
        # set to a torch.distributed group when the batch is sharded across ranks (the launcher does it under torchrun)
        self.shard_group = None
        if os.environ.get("TURBOAE_B200_SHARD") == "1" and torch.distributed.is_available() and torch.distributed.is_initialized():
            self.shard_group = torch.distributed.group.WORLD
        #: 'f16x3' (split-operand tcgen05 kernel, tae_x3.cu: elementwise parity <= 1e-4 at tensor-core speed), 'fp32' (CUDA-core
        #: path, same parity, ~8x slower), 'bf16' (the decoder's fused tcgen05 kernel with the three branches as three conv stacks:
        #: fastest, codes within bf16 rounding of the reference's) or 'auto' (default): 'f16x3' where that kernel covers the
        #: configuration (kernel size 5, <= 104 units, block length <= 510), else 'fp32' -- both meet the same tolerance
        self.precision = getattr(args, "tae_enc_precision", None) or os.environ.get("TURBOAE_B200_ENC_PRECISION", "auto")
        #: training (autograd) path: 'fp32' (CUDA-core kernels) or 'bf16' (tensor cores, train_tc.py)
        from . import train_tc
        self.train_precision = (getattr(args, "tae_train_precision", None) or os.environ.get("TURBOAE_B200_TRAIN_PRECISION")
                                or ("bf16" if (train_tc.supported(args, "enc") and not self.dense) else "fp32"))
        if self.dense:
            self.train_precision = "fp32"

    def set_interleaver(self, p_array):
        self.interleaver.set_parray(p_array)

    def set_parallel(self):
        self._drop_ordered()
        for n in ("enc_cnn_1", "enc_cnn_2", "enc_cnn_3", "enc_linear_1", "enc_linear_2", "enc_linear_3"):
            m = getattr(self, n)
            if not isinstance(m, ParallelShim):
                setattr(self, n, ParallelShim(m))

    # -- canonical flat order of include/turboae_b200.h ------------------------------------------------
    def _walk_ordered_parameters(self):
        out = []
        for i in (1, 2, 3):
            for conv in unwrap(getattr(self, "enc_cnn_%d" % i)).cnns:
                out += [conv.weight, conv.bias]
            lin = unwrap(getattr(self, "enc_linear_%d" % i))
            out += [lin.weight, lin.bias]
        return out

    def config(self, block_len):
        a = self.args
        return _lib.TaeEncConfig(block_len, a.enc_num_layer, a.enc_num_unit, a.enc_kernel_size)

    def _check_supported(self):
        a = self.args
        if getattr(a, "enc_act", "elu") != "elu":
            raise NotImplementedError("enc_act=%r: only 'elu' is built (get_args.py:100 'only elu works')" % a.enc_act)
        if getattr(a, "train_channel_mode", "block_norm") not in ("block_norm", "block_norm_ste"):
            raise NotImplementedError("train_channel_mode=%r is not supported by turboae_b200" % a.train_channel_mode)

    def resolved_precision(self, block_len):
        """The inference path a forward at this block length takes ('auto' resolved; see ``precision``)."""
        if self.precision not in ("auto", "fp32", "bf16", "f16x3"):
            raise _lib.TaeError("encoder precision must be 'auto', 'f16x3', 'fp32' or 'bf16', got %r" % (self.precision,))
        if self.precision != "auto":
            return self.precision
        if _lib.load().tae_enc_packed_bytes_x3(self.config(block_len)):
            return "f16x3"
        if not getattr(self, "_warned_fp32", False):          # said once: the choice is by shape, never silent
            import warnings
            warnings.warn("turboae_b200.ENC_interCNN: the split-operand tensor kernel does not cover this configuration (%s); "
                          "using the fp32 CUDA-core kernels (same tolerance, ~10x slower)" % _lib.load().tae_last_error().decode())
            self._warned_fp32 = True
        return "fp32"

    def encode_unnormalised(self, inputs, stats):
        """x_tx (B, L, 3) before power_constraint; adds (sum, sumsq, count) into the 3 device doubles `stats`."""
        lib = _lib.load()
        B, L = inputs.shape[0], inputs.shape[1]
        dev = inputs.device
        cfg = self.config(L)
        flat = self._flat.get(self.ordered_parameters())
        if flat.device != dev:
            raise _lib.TaeError("encoder parameters are on %s but the input is on %s" % (flat.device, dev))
        perm, inv = self.interleaver.device_index(dev)
        x_tx = torch.empty((B, L, 3), dtype=torch.float32, device=dev)
        precision = self.resolved_precision(L)
        if precision == "f16x3":
            with torch.cuda.device(dev):
                packed = self._flat.derived.get("f16x3")
                if packed is None:
                    nbytes = lib.tae_enc_packed_bytes_x3(cfg)
                    if nbytes == 0:
                        raise _lib.TaeError("f16x3 encoder path unavailable for this configuration (%s); set precision='fp32'"
                                            % lib.tae_last_error().decode())
                    packed = torch.empty(nbytes, dtype=torch.uint8, device=dev)
                    _lib.check(lib.tae_enc_pack_f16x3(cfg, _lib.ptr(flat), _lib.ptr(packed), _lib.stream_ptr(dev)))
                    self._flat.derived["f16x3"] = packed
                ws = self._ws.get(256, dev)
                _lib.check(lib.tae_enc_forward_f16x3(cfg, _lib.ptr(flat), _lib.ptr(packed), _lib.ptr(inputs), _lib.ptr(perm),
                                                      _lib.ptr(inv), _lib.ptr(x_tx), _lib.ptr(stats), B, _lib.ptr(ws), ws.numel(),
                                                      _lib.stream_ptr(dev)))
            return x_tx
        if precision == "bf16":
            with torch.cuda.device(dev):
                packed = self._flat.derived.get("bf16")
                if packed is None:
                    nbytes = lib.tae_enc_packed_bytes(cfg)
                    if nbytes == 0:
                        raise _lib.TaeError("bf16 encoder path unavailable for this configuration (%s); set precision='fp32'"
                                            % lib.tae_last_error().decode())
                    packed = torch.empty(nbytes, dtype=torch.uint8, device=dev)
                    _lib.check(lib.tae_enc_pack_bf16(cfg, _lib.ptr(flat), _lib.ptr(packed), _lib.stream_ptr(dev)))
                    self._flat.derived["bf16"] = packed
                ws = self._ws.get(256, dev)
                _lib.check(lib.tae_enc_forward_bf16(cfg, _lib.ptr(packed), _lib.ptr(inputs), _lib.ptr(perm), _lib.ptr(inv),
                                                    _lib.ptr(x_tx), _lib.ptr(stats), B, _lib.ptr(ws), ws.numel(),
                                                    _lib.stream_ptr(dev)))
            return x_tx
        ws_bytes = lib.tae_enc_workspace_bytes(cfg, B)
        ws = self._ws.get(ws_bytes, dev)
        with torch.cuda.device(dev):
            _lib.check(lib.tae_enc_forward(cfg, _lib.ptr(flat), _lib.ptr(inputs), _lib.ptr(perm), _lib.ptr(x_tx),
                                           _lib.ptr(stats), B, _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)))
        return x_tx

    def _forward_train(self, u):
        """Autograd path (reference encoders.py:362-375 under trainer.py:74).  'bf16': the three branches incl. Linear + ELU in one
        fused forward / backward on the tensor cores (train_tc.EncoderTrainFn); 'fp32' (and the dense variant): the conv stacks
        through the fp32 kernels with torch glue for the 100->1 Linear, ELU and concat.  The power constraint is shard.PowerNorm
        (this package's kernels on the device; statistics and gradient sums all-reduced when the batch is sharded across ranks)."""
        own_stats = None                      # (sum, sum of squares, count) of this rank's x_tx when a kernel already delivered them
        if self.train_precision == "bf16":
            from . import train_tc
            x_tx, own_stats = train_tc.encoder_branches_train(self, u)
        elif self.train_precision == "fp32":
            x = 2.0 * u - 1.0
            outs = []
            for i, inp in ((1, x), (2, x), (3, self.interleaver(x))):
                h = getattr(self, "enc_cnn_%d" % i)(inp)
                outs.append(torch.nn.functional.elu(getattr(self, "enc_linear_%d" % i)(h)))
            x_tx = torch.cat(outs, dim=2)
        else:
            raise _lib.TaeError("train_precision must be 'bf16' or 'fp32', got %r" % (self.train_precision,))
        if self.args.no_code_norm:
            return x_tx
        self._check_supported()
        if getattr(self.args, "precompute_norm_stats", False):
            # encoders.py:110-114 under autograd: the running scalars carry the graph of every batch they have seen in the
            # reference too; here they are treated as constants of the current step (trainer.test is their only user)
            xd = x_tx.detach().double()
            stats = torch.stack([xd.sum(), (xd * xd).sum(), torch.full((), float(x_tx.numel()), dtype=torch.float64, device=x_tx.device)])
            if self.shard_group is not None:
                shard.merge_power_stats(stats, self.shard_group)
            given = self._running_stats(stats)
            codes = (x_tx - given[0]) / given[1]
        else:
            # statistics are merged across ranks only when the batch is sharded (shard_group set); otherwise this call stays local
            codes = shard.PowerNorm.apply(x_tx, self.shard_group if self.shard_group is not None else shard.LOCAL, own_stats)
        if getattr(self.args, "train_channel_mode", "block_norm") == "block_norm_ste":
            codes = STEQuantize.apply(codes, self.args)
        if self.args.enc_truncate_limit > 0:
            codes = torch.clamp(codes, -self.args.enc_truncate_limit, self.args.enc_truncate_limit)
        return codes

    def _variable_block_len(self, block_len):
        """reference encoders.py:353-360: with -is_variable_block_len the interleaver is re-drawn for the length of THIS batch
        (same draw from numpy's global generator as the reference, so a seeded run sees the same permutations)."""
        a = self.args
        if getattr(a, "is_variable_block_len", False) and a.is_interleave != 0:
            import numpy as np
            seed = np.random.randint(0, a.is_interleave)
            self.set_interleaver(np.random.mtrand.RandomState(seed).permutation(np.arange(block_len)))

    def _running_stats(self, stats):
        """reference encoders.py:110-114 (-precompute_norm_stats): running averages of the batch mean / unbiased std over the
        calls so far; returns the 2 device floats (mean, std) to normalise with.  Tiny device-side arithmetic on 1-element
        tensors: no host synchronisation."""
        n = stats[2]
        mean = stats[0] / n
        std = torch.sqrt(torch.clamp((stats[1] - n * mean * mean) / (n - 1.0), min=0.0))
        self.num_test_block += 1.0
        k = self.num_test_block
        self.mean_scalar = (self.mean_scalar * (k - 1.0) + mean.float()) / k
        self.std_scalar = (self.std_scalar * (k - 1.0) + std.float()) / k
        return torch.cat([self.mean_scalar.reshape(1), self.std_scalar.reshape(1)]).contiguous()

    def forward(self, inputs):
        self._check_supported()
        if self.this_device.type != "cuda":
            raise _lib.TaeError("no CUDA device: turboae_b200 has no CPU fallback")
        self._variable_block_len(inputs.shape[1])
        x = inputs.to(device=self.this_device, dtype=torch.float32).contiguous()
        if self.dense or (torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))):
            # (the dense variant: the branches spelled out with the module pieces on the fp32 kernels, with or without autograd)
            return self._forward_train(x)
        if x.dim() != 3 or x.shape[2] != 1:
            raise _lib.TaeError("ENC_interCNN expects (B, L, 1) bits, got %s" % (tuple(x.shape),))
        lib = _lib.load()
        stats = torch.zeros(3, dtype=torch.float64, device=x.device)
        x_tx = self.encode_unnormalised(x, stats)
        if self.args.no_code_norm:
            return x_tx
        if self.shard_group is not None:
            # power_constraint normalises over the WHOLE batch (encoders.py:107-116): merge the per-rank sums
            shard.merge_power_stats(stats, self.shard_group)
        codes = torch.empty_like(x_tx)
        ste = getattr(self.args, "train_channel_mode", "block_norm") == "block_norm_ste"
        with torch.cuda.device(x.device):
            if getattr(self.args, "precompute_norm_stats", False):                              # encoders.py:110-114
                given = self._running_stats(stats)
                _lib.check(lib.tae_power_norm_given_f32(_lib.ptr(x_tx), _lib.ptr(codes), x_tx.numel(), _lib.ptr(given),
                                                        float(self.args.enc_value_limit) if ste else 1.0,
                                                        float(self.args.enc_quantize_level) if ste else 0.0,
                                                        _lib.stream_ptr(x.device)))
            elif ste:                                                                           # encoders.py:118-120
                _lib.check(lib.tae_power_norm_ste_f32(_lib.ptr(x_tx), _lib.ptr(codes), x_tx.numel(), _lib.ptr(stats), None,
                                                      float(self.args.enc_value_limit), float(self.args.enc_quantize_level),
                                                      _lib.stream_ptr(x.device)))
            else:
                _lib.check(lib.tae_power_norm_f32(_lib.ptr(x_tx), _lib.ptr(codes), x_tx.numel(), _lib.ptr(stats), None,
                                                  _lib.stream_ptr(x.device)))
        if self.args.enc_truncate_limit > 0:
            codes = torch.clamp(codes, -self.args.enc_truncate_limit, self.args.enc_truncate_limit)
        return codes
