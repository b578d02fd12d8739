"""Fused optimiser step + CUDA-graph training step.

`FusedOptimizer` mirrors the torch optimisers the reference's `Optimizer` wrapper configures
(tedeous/optimizers/optimizer.py:44-61: Adam / AdamW / SGD) as ONE kernel launch over all parameter tensors
(`tdb200_optimizer_step`), reading the flat gradient straight from the fused plan's output vector: no `.grad`
tensors, no per-tensor launches, no host round trip.  `TrainStep` captures
    pack -> jet / stencil kernels -> reduction -> [NCCL all-reduce] -> parameter update
as one CUDA graph: `Model.train` (tedeous/model.py:174-191) replays it once per epoch."""
import ctypes as C
from typing import List

import numpy as np
import torch

from .. import _native

_KIND = {'Adam': 0, 'AdamW': 1, 'SGD': 2}


class FusedOptimizer:
    """State (moments, step count, learning rate) lives on the device; `step(flat_grad)` enqueues two launches."""

    def __init__(self, name: str, params: List[torch.Tensor], **hyper):
        if name not in _KIND:
            raise NotImplementedError(f'fused optimiser for {name!r} (available: {sorted(_KIND)})')
        self.name, self.kind = name, _KIND[name]
        self.params = [p for p in params]
        for p in self.params:
            if p.dtype != torch.float32 or not p.is_cuda or not p.is_contiguous():
                raise ValueError('FusedOptimizer needs contiguous float32 CUDA parameters')
        if len(self.params) > 40:
            raise ValueError('too many parameter tensors')
        self.device = self.params[0].device
        self.lib = _native.load()
        unknown = set(hyper) - {'lr', 'betas', 'eps', 'weight_decay', 'momentum', 'dampening', 'nesterov', 'amsgrad',
                                'maximize', 'foreach', 'capturable', 'differentiable', 'fused'}
        if unknown:
            raise ValueError(f'unknown optimiser options {sorted(unknown)}')
        if hyper.get('amsgrad') or hyper.get('maximize') or hyper.get('nesterov') or hyper.get('dampening', 0) != 0:
            raise NotImplementedError('amsgrad / maximize / nesterov / dampening are not provided by the fused optimiser')
        lr = float(hyper.get('lr', 1e-3))
        if self.kind == 2:
            b1, b2, eps = float(hyper.get('momentum', 0.0)), 0.0, 0.0
            wd = float(hyper.get('weight_decay', 0.0))
        else:
            b1, b2 = (float(x) for x in hyper.get('betas', (0.9, 0.999)))
            eps = float(hyper.get('eps', 1e-8))
            wd = float(hyper.get('weight_decay', 1e-2 if self.kind == 1 else 0.0))
        self.n = sum(p.numel() for p in self.params)
        self.hyper = torch.tensor([lr, b1, b2, eps, wd], dtype=torch.float32, device=self.device)
        self.m = torch.zeros(self.n, dtype=torch.float32, device=self.device)
        self.v = torch.zeros(self.n if self.kind != 2 else 1, dtype=torch.float32, device=self.device)
        self.step_count = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._ptrs = (C.c_void_p * len(self.params))(*[p.data_ptr() for p in self.params])
        self._sizes = np.array([p.numel() for p in self.params], dtype=np.int64)

    @property
    def lr(self) -> float:
        return float(self.hyper[0])

    def set_lr(self, lr: float) -> None:
        """Schedulers: written in stream order, so it may be called between graph replays."""
        self.hyper[0:1].fill_(float(lr))

    def step(self, flat_grad: torch.Tensor) -> None:
        if flat_grad.numel() != self.n or flat_grad.dtype != torch.float32:
            raise ValueError(f'gradient has {flat_grad.numel()} entries for {self.n} parameters')
        for i, p in enumerate(self.params):          # the tensors must keep their storage (in-place updates do)
            if p.data_ptr() != self._ptrs[i]:
                self._ptrs[i] = p.data_ptr()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _native.check(self.lib.tdb200_optimizer_step(
            self.kind, len(self.params), self._ptrs, _native.np_ptr(self._sizes), flat_grad.data_ptr(), self.m.data_ptr(),
            self.v.data_ptr(), self.hyper.data_ptr(), self.step_count.data_ptr(), stream), 'tdb200_optimizer_step')


class TrainStep:
    """One training step of a fused `Solution` as a single CUDA-graph replay.

    `step()` -> the static output vector [loss, loss_normalized, slot MSEs ..., gradient ...] of the step just enqueued
    (device memory, no synchronisation; read `out[0]` only when a callback needs the number)."""

    def __init__(self, solution, optimizer: FusedOptimizer, use_graph: bool = True):
        self.sol, self.opt = solution, optimizer
        if solution.tol != 0 or solution.weak_form not in (None, []):
            raise NotImplementedError('graph-captured training step: causal / weak-form losses keep the eager loop')
        self.graph = None
        dev = solution.grid.device
        if use_graph:
            cur = torch.cuda.current_stream(dev)
            side = torch.cuda.Stream(dev)
            side.wait_stream(cur)
            with torch.cuda.stream(side):                # warm-up outside the capture: lazy allocations, NCCL, function attributes
                for _ in range(3):
                    self._enqueue(dry=True)
            cur.wait_stream(side)
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.out = self._enqueue()
        else:
            self.out = None

    def _enqueue(self, dry: bool = False):
        sol = self.sol
        if sol.mode == 'mat':
            out, grad = sol._plan.loss_grad_raw(sol.model)
            flat = grad.reshape(-1)
        else:
            out, flat = sol._run_plan()
        if not dry:
            self.opt.step(flat)
        return out

    def step(self) -> torch.Tensor:
        if self.graph is not None:
            self.graph.replay()
            return self.out
        self.out = self._enqueue()
        return self.out
