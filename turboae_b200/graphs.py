"""CUDA-graph capture of a whole training step.

The reference's trainer (trainer.py:33-76) is a Python loop: per step it draws the batch on the CPU, runs forward, backward
and the optimizer through ~170 small host-side calls (autograd nodes, Adam over 162 parameter tensors).  On a B200 those
~4 ms of host work are longer than the ~3 ms of kernels of a 1000-codeword step, and under data parallelism every rank's host
jitter is exposed at the gradient all-reduce.  A step whose tensors live at fixed addresses can instead be captured ONCE
(kernels of this package, the NCCL all-reduce, a `capturable` optimizer) and replayed: `GraphedStep(fn)()`.

`fn` must be capture-safe: no host synchronisation (.item(), .cpu()), random numbers only from torch's CUDA generator,
`optimizer.zero_grad(set_to_none=True)` inside (the tensor-core autograd path hands the gradients out as views of one
persistent flat buffer), optimizer built with `capturable=True`."""
from __future__ import annotations

import torch


class GraphedStep:
    def __init__(self, fn, warmup: int = 3, device=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.fn = fn
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(warmup):            # allocates every persistent buffer, initialises NCCL, builds the job lists
                fn()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        from . import _flat
        _flat.invalidate_all()                  # the flat copy / bf16 images must be rebuilt INSIDE the captured step
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = fn()

    def __call__(self):
        self.graph.replay()
        return self.out
