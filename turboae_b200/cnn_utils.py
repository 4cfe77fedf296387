"""SameShapeConv1d with the reference's module tree (reference cnn_utils.py:6-46): a ModuleList ``cnns`` of
``torch.nn.Conv1d(k, stride 1, padding k//2)`` so parameter names, shapes and default initialisation are the
reference's; ``forward`` runs each layer as one fused conv + bias + ELU kernel on channel-last (B, L, C)."""
from __future__ import annotations

import torch

from . import _lib


class SameShapeConv1d(torch.nn.Module):
    def __init__(self, num_layer, in_channels, out_channels, kernel_size, activation="elu", no_act=False):
        super().__init__()
        if activation != "elu":
            raise NotImplementedError("turboae_b200.SameShapeConv1d implements activation='elu' only "
                                      "(the only one the hot path uses, cnn_utils.py:24-25)")
        if kernel_size % 2 == 0 or kernel_size > 9:
            raise NotImplementedError("kernel_size must be odd and <= 9 (got %d)" % kernel_size)
        self.cnns = torch.nn.ModuleList()
        self.num_layer = num_layer
        self.no_act = no_act
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, kernel_size
        for idx in range(num_layer):
            self.cnns.append(torch.nn.Conv1d(in_channels=in_channels if idx == 0 else out_channels,
                                             out_channels=out_channels, kernel_size=kernel_size, stride=1,
                                             padding=kernel_size // 2, dilation=1, groups=1, bias=True))

    def forward(self, inputs):
        _lib.require_cuda(inputs, "SameShapeConv1d input")
        x = inputs.to(torch.float32).contiguous()
        params = []
        for conv in self.cnns:
            params += [conv.weight, conv.bias]
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params)):
            return _ConvStackFn.apply(x, not self.no_act, *params)           # training: activations are kept
        return _conv_stack_forward(x, params, not self.no_act, keep=False)[-1]


class DenseSameShapeConv1d(torch.nn.Module):
    """reference cnn_utils.py:49-82 (chosen by decoders.py:173 for every -encoder other than TurboAE_rate3_cnn): layer idx reads
    the concatenation of the stack input and ALL earlier layers' ELU outputs (in_channels + idx * out_channels channels).  Same
    module tree (``cnns``) and parameter shapes as the reference; every layer is the fused conv + bias + ELU kernel (forward) and
    the conv backward kernels (autograd), the concatenation is torch glue on channel-last tensors."""

    def __init__(self, num_layer, in_channels, out_channels, kernel_size):
        super().__init__()
        if kernel_size % 2 == 0 or kernel_size > 9:
            raise NotImplementedError("kernel_size must be odd and <= 9 (got %d)" % kernel_size)
        self.cnns = torch.nn.ModuleList()
        self.num_layer = num_layer
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, kernel_size
        for idx in range(num_layer):
            self.cnns.append(torch.nn.Conv1d(in_channels=in_channels + idx * out_channels, out_channels=out_channels,
                                             kernel_size=kernel_size, stride=1, padding=kernel_size // 2, dilation=1, groups=1,
                                             bias=True))

    def forward(self, inputs):
        _lib.require_cuda(inputs, "DenseSameShapeConv1d input")
        this_input = inputs.to(torch.float32).contiguous()
        output = None
        for idx, conv in enumerate(self.cnns):
            if idx > 0:
                this_input = torch.cat([this_input, output], dim=2).contiguous()        # cnn_utils.py:73 (dim 1 of the NCL view)
            if torch.is_grad_enabled() and (this_input.requires_grad or conv.weight.requires_grad or conv.bias.requires_grad):
                output = _ConvStackFn.apply(this_input, True, conv.weight, conv.bias)
            else:
                output = _conv_stack_forward(this_input, [conv.weight, conv.bias], True, keep=False)[-1]
        return output


def _conv_layer(x, w, b, apply_elu):
    lib = _lib.load()
    B, L, _ = x.shape
    cout, cin, k = w.shape
    ws_bytes = lib.tae_conv1d_workspace_bytes(cin, cout, k)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
    out = torch.empty((B, L, cout), dtype=torch.float32, device=x.device)
    _lib.check(lib.tae_conv1d_elu_f32(_lib.ptr(x), _lib.ptr(out), _lib.ptr(w), _lib.ptr(b), B, L, cin, cout, k,
                                      1 if apply_elu else 0, _lib.ptr(ws), ws_bytes, _lib.stream_ptr(x.device)))
    return out


def _conv_stack_forward(x, params, apply_elu, keep):
    """[x, h_1, ..., h_n] (keep) or [h_n]: every layer is one fused conv + bias + ELU kernel (cnn_utils.py:36-46)."""
    acts = [x]
    with torch.cuda.device(x.device):
        for j in range(0, len(params), 2):
            h = _conv_layer(acts[-1], params[j].detach().contiguous(), params[j + 1].detach().contiguous(), apply_elu)
            if keep:
                acts.append(h)
            else:
                acts = [h]
    return acts


class _ConvStackFn(torch.autograd.Function):
    """Training path of SameShapeConv1d (reference trainer.py:74 backpropagates through cnn_utils.py:36-46): forward keeps
    every layer's output; backward runs, per layer, the weight-gradient kernel and the same conv kernel with transposed,
    tap-flipped weights on g = dy * ELU'(z) (``tae_conv1d_elu_bwd_f32``)."""

    @staticmethod
    def forward(ctx, x, apply_elu, *params):
        acts = _conv_stack_forward(x, params, apply_elu, keep=True)
        ctx.apply_elu = apply_elu
        ctx.save_for_backward(*acts, *[p for p in params[0::2]])
        ctx.n_layer = len(params) // 2
        ctx.needs = [x.requires_grad] + [p.requires_grad for p in params]
        return acts[-1]

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        n = ctx.n_layer
        saved = ctx.saved_tensors
        acts, weights = saved[:n + 1], saved[n + 1:]
        dy = dy.contiguous()
        grads = [None] * (2 * n)
        dev = dy.device
        with torch.cuda.device(dev):
            for j in range(n - 1, -1, -1):
                w = weights[j].detach().contiguous()
                cout, cin, k = w.shape
                x, y = acts[j], acts[j + 1]
                B, L, _ = x.shape
                need_dx = j > 0 or ctx.needs[0]
                need_dw = ctx.needs[1 + 2 * j] or ctx.needs[2 + 2 * j]
                dx = torch.empty_like(x) if need_dx else None
                dw = torch.zeros_like(w) if need_dw else None
                db = torch.zeros(cout, dtype=torch.float32, device=dev) if need_dw else None
                ws_bytes = lib.tae_conv1d_bwd_workspace_bytes(cin, cout, k)
                ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
                if need_dx or need_dw:
                    _lib.check(lib.tae_conv1d_elu_bwd_f32(_lib.ptr(x), _lib.ptr(y), _lib.ptr(dy), _lib.ptr(w), _lib.ptr(dx),
                                                          _lib.ptr(dw), _lib.ptr(db), B, L, cin, cout, k,
                                                          1 if ctx.apply_elu else 0, _lib.ptr(ws), ws_bytes,
                                                          _lib.stream_ptr(dev)))
                grads[2 * j], grads[2 * j + 1] = dw, db
                dy = dx
        return (dy if ctx.needs[0] else None, None, *grads)
