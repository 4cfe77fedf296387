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
        if torch.is_grad_enabled() and (inputs.requires_grad or any(p.requires_grad for p in self.parameters())):
            raise NotImplementedError("turboae_b200: backward of the conv stack is not built yet "
                                      "(SURVEY.md section 8(f) row 1); wrap inference in torch.no_grad()")
        lib = _lib.load()
        x = inputs.to(torch.float32).contiguous()
        B, L, _ = x.shape
        with torch.cuda.device(x.device):
            for conv in self.cnns:
                cout, cin, k = conv.weight.shape
                ws_bytes = lib.tae_conv1d_workspace_bytes(cin, cout, k)
                ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
                out = torch.empty((B, L, cout), dtype=torch.float32, device=x.device)
                _lib.check(lib.tae_conv1d_elu_f32(_lib.ptr(x), _lib.ptr(out), _lib.ptr(conv.weight.detach().contiguous()),
                                                  _lib.ptr(conv.bias.detach().contiguous()), B, L, cin, cout, k,
                                                  0 if self.no_act else 1, _lib.ptr(ws), ws_bytes,
                                                  _lib.stream_ptr(x.device)))
                x = out
        return x
