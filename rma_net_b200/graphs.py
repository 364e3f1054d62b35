"""CUDA-graph replay of a whole loss step.

The reference trains at B=4, N=2048, T=7 stages (train_sample.py:15-30): one `Loss.forward` + `backward` is a few dozen
launches of 10-50 us of device work in total, so an eager step is bound by the host (Python + launch latency), not by
the kernels.  `GraphedStep` captures `loss = fn(*inputs); loss.backward()` once and replays it: one graph launch per
step, the gradients land in static tensors.  All ops of this package are capture-safe (caller-provided streams, no host
synchronisation, workspaces from torch's allocator).

    step = GraphedStep(lambda d0, d1, tgt: loss_module(src, tgt, [d0, d1], ...)[0], [d0, d1, tgt], grads_for=[0, 1])
    loss, (g0, g1) = step(new_d0, new_d1, new_tgt)      # copies the new values in, replays, returns static tensors
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch


class GraphedStep:
    """Capture `fn(*inputs)` (a scalar loss) and its backward w.r.t. the inputs listed in `grads_for`.

    inputs      example CUDA tensors; their shapes / dtypes / strides are frozen, their values are copied in per call
    grads_for   indices of the inputs that need a gradient (default: every floating-point input)
    warmup      eager runs on a side stream before the capture (allocator warm-up, lazy initialisation)
    The returned loss and gradient tensors are static: they are overwritten by the next call."""

    def __init__(self, fn: Callable[..., torch.Tensor], inputs: Sequence[torch.Tensor],
                 grads_for: Optional[Sequence[int]] = None, warmup: int = 3):
        if not inputs or not all(isinstance(t, torch.Tensor) and t.is_cuda for t in inputs):
            raise RuntimeError("GraphedStep needs CUDA tensors as example inputs")
        if grads_for is None:
            grads_for = [i for i, t in enumerate(inputs) if t.is_floating_point()]
        self._grads_for = list(grads_for)
        self._static = [t.detach().clone(memory_format=torch.preserve_format) for t in inputs]
        for i in self._grads_for:
            self._static[i].requires_grad_(True)
        self._fn = fn
        dev = inputs[0].device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._run_eager()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._loss, self._grads = self._run_eager()

    def _run_eager(self) -> Tuple[torch.Tensor, List[torch.Tensor]]:
        for i in self._grads_for:
            self._static[i].grad = None
        loss = self._fn(*self._static)
        if loss.dim() != 0:
            raise RuntimeError("GraphedStep: fn must return a scalar loss")
        grads = torch.autograd.grad(loss, [self._static[i] for i in self._grads_for], allow_unused=True)
        return loss.detach(), [g if g is not None else torch.zeros_like(self._static[i])
                               for g, i in zip(grads, self._grads_for)]

    def __call__(self, *inputs: torch.Tensor) -> Tuple[torch.Tensor, List[torch.Tensor]]:
        if len(inputs) != len(self._static):
            raise RuntimeError(f"GraphedStep: expected {len(self._static)} inputs, got {len(inputs)}")
        for s, t in zip(self._static, inputs):
            if t.shape != s.shape or t.dtype != s.dtype:
                raise RuntimeError(f"GraphedStep: input {tuple(t.shape)} {t.dtype} does not match the captured "
                                   f"{tuple(s.shape)} {s.dtype}")
            if t.data_ptr() != s.data_ptr():
                with torch.no_grad():
                    s.copy_(t, non_blocking=True)
        self._graph.replay()
        return self._loss, self._grads

    def replay(self) -> Tuple[torch.Tensor, List[torch.Tensor]]:
        """Replay on the current contents of the static inputs (`static_inputs` may be written in place)."""
        self._graph.replay()
        return self._loss, self._grads

    @property
    def static_inputs(self) -> List[torch.Tensor]:
        return self._static
