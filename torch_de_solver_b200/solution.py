"""`Solution` - the drop-in boundary (tedeous/solution.py:23-168).

`Solution(grid, equal_cls, model, mode, weak_form, lambda_operator, lambda_bound, tol, derivative_points,
batch_size).evaluate()` returns `(loss [1], loss_normalized [1])` exactly like the reference, but one call
enqueues the fused CUDA plan (libtedeous_b200.so) instead of walking the operator dicts through ATen:

* loss and the full parameter gradient are produced by the same launch sequence; `loss.backward()` (or
  `torch.autograd.grad(loss, params)`) just hands the already-computed gradient to autograd through a
  custom `torch.autograd.Function`, so `Closure._closure` (optimizers/closure.py:49-64) works unchanged;
* `op`, `bval`, `true_bval`, `bval_keys`, `bval_length` (read by callbacks, adaptive_lambda.py:91-105) are
  materialised lazily by a forward-only launch (`tdb200_eval_fields`);
* there is no CPU path: a CPU default device raises.
"""
from __future__ import annotations

import ctypes as C
from copy import deepcopy
from typing import List, Optional, Tuple

import numpy as np
import torch

from . import _native
from .device import check_device, device_type
from .input_preprocessing import lambda_prepare
from .plan import (ProblemIR, UnsupportedProblem, flatten, lower_problem)


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f'{what} lives on {t.device}: the tedeous-b200 hot path runs on CUDA (sm_100a) only '
                           "and has no CPU fallback - call solver_device('cuda') first")


def _dist_ready() -> bool:
    """A sharded plan built outside a torch.distributed job (tests that add the shards up by hand) has no collective."""
    import torch.distributed as dist
    return dist.is_available() and dist.is_initialized()


_LIB_COMMS = {}          # process group -> communicator handle; kept for the life of the process (never destroyed at exit)


def library_comm(lib, device, rank: int, world: int, group=None):
    import torch.distributed as dist
    key = (id(group) if group is not None else 0, device.index)
    if key not in _LIB_COMMS:
        uid = torch.zeros(128, dtype=torch.uint8, device='cpu')
        if rank == 0:
            buf = (C.c_char * 128)()
            _native.check(lib.tdb200_comm_unique_id(buf), 'tdb200_comm_unique_id')
            uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        uid = uid.to(device)
        dist.broadcast(uid, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        raw = bytes(uid.cpu().numpy().tobytes())
        handle = C.c_void_p()
        dev_index = device.index if device.index is not None else torch.cuda.current_device()
        _native.check(lib.tdb200_comm_create(raw, rank, world, dev_index, C.byref(handle)), 'tdb200_comm_create')
        _LIB_COMMS[key] = handle
    return _LIB_COMMS[key]


class FusedPlan:
    """Owns one tdb200_plan (modes 'NN' / 'autograd') and the device buffers bound to it."""

    def __init__(self, ir: ProblemIR, device: torch.device, impl: int = 0):
        self.ir = ir
        self.device = device
        self.lib = _native.load()
        self.flat = flatten(ir, device)
        f = self.flat
        net = _native.NetDesc()
        net.n_layers = ir.net.n_layers
        for i, w in enumerate(ir.net.widths):
            net.widths[i] = w
        net.n_cparams = len(ir.net.coeff_params)
        handle = C.c_void_p()
        dev_index = device.index if device.index is not None else torch.cuda.current_device()
        _native.check(self.lib.tdb200_plan_create(
            C.byref(net), len(f.seg), _native.np_ptr(f.seg), len(f.terms), _native.np_ptr(f.terms),
            len(f.factors), _native.np_ptr(f.factors), len(f.comb), _native.np_ptr(f.comb), ir.n_slots,
            dev_index, C.byref(handle)), 'tdb200_plan_create')
        self.handle = handle
        _native.check(self.lib.tdb200_plan_set_points(
            handle, f.points.data_ptr(), f.points.shape[0], f.targets.data_ptr(), f.targets.numel(),
            f.coeffs.data_ptr(), f.coeffs.numel()), 'tdb200_plan_set_points')
        self.set_lambdas(ir.slot_lambda)
        if impl:
            _native.check(self.lib.tdb200_plan_set_impl(handle, impl), 'tdb200_plan_set_impl')
        self.out_size = int(self.lib.tdb200_plan_out_size(handle))
        self.n_params = int(self.lib.tdb200_plan_n_params(handle))
        self.n_fields = int(self.lib.tdb200_plan_n_fields(handle))
        self.launches_per_call = int(self.lib.tdb200_plan_launches_per_call(handle))
        self._ptrs = (C.c_void_p * (2 * ir.net.n_layers + len(ir.net.coeff_params)))()

    @property
    def slot_lambda(self):
        return self.ir.slot_lambda

    def set_lambdas(self, slot_lambda):
        self.ir.slot_lambda = [float(x) for x in slot_lambda]
        lam = np.asarray(self.ir.slot_lambda, dtype=np.float64)
        ln = np.asarray(self.ir.slot_len, dtype=np.float64)
        _native.check(self.lib.tdb200_plan_set_slots(self.handle, _native.np_ptr(lam), _native.np_ptr(ln)),
                      'tdb200_plan_set_slots')

    def set_row_weights(self, w: Optional[torch.Tensor]):
        """Per-row loss weights of the interior operator rows (causal loss); None switches them off."""
        self._row_w = None if w is None else w.detach().to(self.device, torch.float32).contiguous()
        if self._row_w is not None and self._row_w.numel() != self.ir.segments[0].n_groups:
            raise ValueError('one weight per interior operator row expected')
        _native.check(self.lib.tdb200_plan_set_row_weights(
            self.handle, None if self._row_w is None else self._row_w.data_ptr()), 'tdb200_plan_set_row_weights')

    def vjp(self, seeds: torch.Tensor) -> torch.Tensor:
        """Vector-Jacobian product: sum_fields seeds * d field / d theta -> flat [n_params] (parameter order of the
        gradient).  `seeds` has the layout of `eval_fields()[0]`; one fused launch in seed mode."""
        seeds = seeds.detach().to(self.device, torch.float32).contiguous()
        if seeds.numel() != max(self.n_fields, 1):
            raise ValueError(f'expected {self.n_fields} field cotangents, got {seeds.numel()}')
        _native.check(self.lib.tdb200_plan_set_field_seeds(self.handle, seeds.data_ptr()), 'tdb200_plan_set_field_seeds')
        try:
            out = self.loss_grad()
        finally:
            _native.check(self.lib.tdb200_plan_set_field_seeds(self.handle, None), 'tdb200_plan_set_field_seeds')
        return out[self.out_size - self.n_params:]

    def refresh_coeffs(self):
        """Re-evaluates every callable coefficient into its per-row buffer (the reference calls coeff(grid) on every
        step, tedeous/derivative.py:41-42, 114-115; here they are evaluated at lowering and on request)."""
        with torch.no_grad():
            for off, n, fn, rows, sl, _seg in self.flat.coeff_fns or ():
                vals = fn(rows).reshape(-1).detach()
                if sl is not None:
                    vals = vals[sl[0]:sl[1]]
                self.flat.coeffs[off:off + n].copy_(vals.to(self.flat.coeffs.dtype))

    @property
    def kernel_path(self) -> str:
        """Kernels serving the interior segment under the current impl setting."""
        return {1: 'simt-fp32', 2: 'tcgen05-3xtf32 (dW in TMEM)', 3: 'tcgen05-3xtf32 streamed (jet_tcs + wgrad_gemm)'}[
            int(self.lib.tdb200_plan_kernel_path(self.handle))]

    def comm_init(self, rank: int, world: int, group=None):
        """Lends the plan this process's NCCL communicator under the C ABI (made once per process group: the 128-byte
        unique id comes from rank 0 through torch.distributed); afterwards every loss_grad call all-reduces its output
        inside the library."""
        import os
        # peer-memory all-reduce: opt-in (TDB200_COLLECTIVE=peer).  Measured on 2 GPUs, BASELINE config 1: 0.134 ms per
        # step against 0.136 ms with ncclAllReduce - the cost of the step over several ranks is the rendezvous itself
        if os.environ.get('TDB200_COLLECTIVE', 'nccl') == 'peer' and self._peer_init(rank, world, group):
            self.has_comm = True
            return
        _native.check(self.lib.tdb200_plan_set_comm(self.handle, library_comm(self.lib, self.device, rank, world, group)),
                      'tdb200_plan_set_comm')
        self.has_comm = True

    def _peer_init(self, rank: int, world: int, group=None) -> bool:
        """Ranks = GPUs of one box: the all-reduce of the [loss terms | gradient] vector runs over CUDA-IPC mapped peer
        memory (csrc/peer.cu, `tdb200_peer_allreduce_vec`: every rank pushes its vector into every rank's inbox and adds
        the rows in rank order) - one small kernel of the library behind the reduction kernel instead of ncclAllReduce.
        False (nothing changed) when the blocks cannot be mapped, e.g. ranks on different hosts."""
        from .mat import open_peer
        handle = C.c_void_p()
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        cap = (self.out_size + 3) // 4 * 4
        created = self.lib.tdb200_peer_create(rank, world, 0, cap, dev_index, C.byref(handle)) == 0
        if not open_peer(self.lib, handle, created, world, group):
            return False
        self._peer = handle
        _native.check(self.lib.tdb200_plan_set_peer(self.handle, handle), 'tdb200_plan_set_peer')
        return True

    has_comm = False
    _peer = None

    def set_interior_rows(self, pts: torch.Tensor, n_valid: int):
        """Mini-batching (tedeous/eval.py:124-141, 174-182): replace the points of the interior segment by `pts` (as many rows
        as the plan was built with; rows >= n_valid are padding with loss weight 0) and re-evaluate its callable
        coefficients on them; the loss divides by n_valid."""
        seg = self.ir.segments[0]
        if pts.shape[0] != seg.n_groups:
            raise ValueError('batch must have the plan\'s number of interior rows')
        with torch.no_grad():
            self.flat.points[:seg.n_groups].copy_(pts.to(self.flat.points.dtype))
            for off, n, fn, rows, sl, sidx in self.flat.coeff_fns or ():
                if sidx == 0:
                    self.flat.coeffs[off:off + n].copy_(fn(pts).reshape(-1).detach().to(self.flat.coeffs.dtype))
        n_eq = len(seg.cols)
        if self.ir.slot_len[0] != n_valid:
            self.ir.slot_len = [n_valid] * n_eq + list(self.ir.slot_len[n_eq:])
            self.set_lambdas(self.ir.slot_lambda)
            if n_valid < seg.n_groups:
                w = torch.zeros(seg.n_groups, dtype=torch.float32, device=self.device)
                w[:n_valid] = 1.0
                self.set_row_weights(w)
            else:
                self.set_row_weights(None)

    def set_impl(self, impl: int):
        _native.check(self.lib.tdb200_plan_set_impl(self.handle, impl), 'tdb200_plan_set_impl')
        self.launches_per_call = int(self.lib.tdb200_plan_launches_per_call(self.handle))

    def _param_ptrs(self):
        for i, p in enumerate(self.ir.net.param_tensors()):
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError('network parameters must be contiguous float32 CUDA tensors '
                                   f'(got {p.dtype} on {p.device})')
            self._ptrs[i] = p.data_ptr()
        return self._ptrs

    def loss_grad(self) -> torch.Tensor:
        """-> flat [2 + n_slots + n_params] tensor (loss, loss_normalized, slot MSEs, gradient)."""
        out = torch.empty(self.out_size + 3, dtype=torch.float32, device=self.device)[:self.out_size]   # (float4 tail of the peer all-reduce)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _native.check(self.lib.tdb200_loss_grad(self.handle, self._param_ptrs(), out.data_ptr(), stream),
                      'tdb200_loss_grad')
        return out

    def capture(self):
        """CUDA graph of one loss + gradient call (pack, interior and boundary launches with their fork / join, reduction)
        for steps that are launch bound (BASELINE config 1: ~0.15 ms of GPU work in 8 launches).  -> (replay callable,
        out): `out` is a static [2 + n_slots + n_params] tensor refreshed by every replay; the parameter tensors must
        keep their storage (in-place optimiser updates do)."""
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(3):                          # warm-up outside the capture (function attributes, side stream)
                self.loss_grad()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = self.loss_grad()
        return graph.replay, out

    def eval_fields(self) -> Tuple[torch.Tensor, torch.Tensor]:
        fields = torch.empty(max(self.n_fields, 1), dtype=torch.float32, device=self.device)
        out = torch.empty(self.out_size + 3, dtype=torch.float32, device=self.device)[:self.out_size]   # (float4 tail of the peer all-reduce)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _native.check(self.lib.tdb200_eval_fields(self.handle, self._param_ptrs(), fields.data_ptr(),
                                                  out.data_ptr(), stream), 'tdb200_eval_fields')
        return fields, out

    def jacobian_rows(self, segment: int, col: int = 0) -> torch.Tensor:
        """Per-residual Jacobian of one segment's residual column: [n_groups, n_params], row r = d field[r, col] / d theta
        in the parameter order of the flat gradient (tdb200_jacobian_rows; the reference's NGD obtains these rows with
        one torch.autograd.grad per residual, tedeous/optimizers/ngd.py:57-77)."""
        n = self.ir.segments[segment].n_groups
        pad = int(self.lib.tdb200_plan_n_params_pad(self.handle))
        rows = torch.empty(max(n, 1), pad, dtype=torch.float32, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _native.check(self.lib.tdb200_jacobian_rows(self.handle, self._param_ptrs(), segment, col, rows.data_ptr(),
                                                    stream), 'tdb200_jacobian_rows')
        return rows[:n, :self.n_params]

    def __del__(self):
        try:
            if getattr(self, 'handle', None):
                self.lib.tdb200_plan_destroy(self.handle)
                self.handle = None
            if getattr(self, '_peer', None):
                self.lib.tdb200_peer_destroy(self._peer)
                self._peer = None
        except Exception:
            pass


def causal_row_weights(res_rows: torch.Tensor, tol: float, n_t: int, n_global: int, shard_range=None, world: int = 1,
                       group=None) -> torch.Tensor:
    """No-grad weights of the causal loss (tedeous/losses.py:160-171) for this rank's block of interior rows.

    res_rows [n_local] = sum over equations of op^2 for the rows [lo, hi) = `shard_range` of the global row order (grid
    column 0 slowest).  With M = n_global / n_t rows per time slice, w[t, j] = exp(-tol * sum_{s < t} res[s, j]): an
    exclusive prefix sum over the time slices, column by column, instead of the reference's [n_t, n_t] triangular
    matmul.  Several ranks own whole time slices each; the prefix of a rank starts from the column totals of the ranks
    below it: one all-gather of M floats per rank."""
    if n_global % n_t:
        n_t = n_global                                     # the reference's fallback (losses.py:162-165): one row per slice
    m = n_global // n_t
    lo, hi = shard_range if shard_range is not None else (0, res_rows.numel())
    if world > 1 and (lo % m or hi % m):
        raise UnsupportedProblem(f'causal loss over {world} ranks: rows [{lo}, {hi}) of this rank split a time slice of '
                                 f'{m} rows; choose a grid whose number of time slices ({n_t}) is a multiple of the ranks')
    res = res_rows.detach().reshape(-1, m)
    excl = torch.cumsum(res, 0) - res
    if world > 1:
        import torch.distributed as dist
        tot = res.sum(0).contiguous()
        parts = [torch.empty_like(tot) for _ in range(world)]
        dist.all_gather(parts, tot, group=group)
        rank = dist.get_rank(group)
        if rank > 0:
            excl = excl + torch.stack(parts[:rank]).sum(0)
    return torch.exp(-float(tol) * excl).reshape(-1)


class _FusedLoss(torch.autograd.Function):
    """loss = plan(params); backward returns the gradient the same launch already produced."""

    @staticmethod
    def forward(ctx, sol, *params):
        out, flat_grad = sol._run_plan()
        ctx.flat_grad = flat_grad
        ctx.shapes = [p.shape for p in params]
        sol._last_out = out
        return out[0:1].clone()

    @staticmethod
    @torch.autograd.function.once_differentiable        # the gradient is a constant of the launch: no double backward
    def backward(ctx, g):
        flat = ctx.flat_grad.reshape(-1) * g
        grads, off = [], 0
        for shp in ctx.shapes:
            n = int(np.prod(shp)) if len(shp) else 1
            grads.append(flat[off:off + n].view(shp))
            off += n
        return (None, *grads)


class _FusedFields(torch.autograd.Function):
    """fields = plan.eval_fields(params) (flat, every per-point operator / boundary value); backward = one fused launch
    in vector-Jacobian mode.  Makes any torch function of the per-point fields differentiable w.r.t. the network."""

    @staticmethod
    def forward(ctx, plan, *params):
        fields, _ = plan.eval_fields()
        ctx.plan = plan
        ctx.shapes = [p.shape for p in params]
        return fields

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        flat = ctx.plan.vjp(g)
        grads, off = [], 0
        for shp in ctx.shapes:
            n = int(np.prod(shp)) if len(shp) else 1
            grads.append(flat[off:off + n].view(shp))
            off += n
        return (None, *grads)


class Solution:
    """Same constructor and public attributes as tedeous.solution.Solution."""

    def __init__(self, grid: torch.Tensor, equal_cls, model, mode: str, weak_form, lambda_operator,
                 lambda_bound, tol: float = 0, derivative_points: int = 2, batch_size: int = None,
                 shard: Optional[Tuple[int, int]] = None, process_group=None, nn_interior: str = 'jet',
                 impl: int = 0, callable_coeffs: str = 'once', batch_generator: torch.Generator = None,
                 collective: str = 'library'):
        """Extensions after `batch_size`: shard / process_group (multi-GPU), nn_interior ('jet' | 'literal'), impl (kernel
        choice), callable_coeffs: 'once' - callable coefficients are evaluated at lowering (and by `refresh_coeffs()`),
        'every_step' - re-evaluated by every `evaluate()` like the reference does; batch_generator: the generator of the
        mini-batch shuffle (default, as in the reference: a fresh torch.Generator on the grid's device)."""
        if collective not in ('library', 'torch'):
            raise ValueError("collective must be 'library' (NCCL inside libtedeous_b200.so) or 'torch' (torch.distributed)")
        self._collective = collective if torch.cuda.is_available() else 'torch'
        if callable_coeffs not in ('once', 'every_step'):
            raise ValueError("callable_coeffs must be 'once' or 'every_step'")
        self._callable_coeffs = callable_coeffs
        weak = weak_form not in (None, [])
        if weak and mode == 'mat':
            raise UnsupportedProblem('weak-form loss is implemented for modes NN / autograd only (SURVEY 8 a11)')
        if weak and (tol != 0 or (shard is not None and shard[1] > 1)):
            raise UnsupportedProblem('weak-form loss: no causal weights, not sharded over ranks')
        if tol != 0 and mode == 'mat' and shard is not None and shard[1] > 1:
            raise UnsupportedProblem('causal loss (tol != 0) in mat mode is not sharded over ranks')
        self.grid = check_device(grid)
        # Mini-batches (tedeous/eval.py:124-141, 174-182; solution.py:159-166): mode 'autograd' draws shuffled batches of the
        # grid rows for the operator; in mode 'NN' the reference's batches never reach the operator (Derivative_NN ignores
        # grid_points), so batch_size changes nothing there; mode 'mat' has no row batches to draw.
        self._batching = False
        if batch_size is not None:
            if mode == 'mat':
                raise UnsupportedProblem("batch_size in mode 'mat' (the reference would split the [d, N0, N1] grid tensor along "
                                         "its coordinate axis)")
            if weak or tol != 0 or (shard is not None and shard[1] > 1):
                raise UnsupportedProblem('mini-batches with the weak-form / causal loss or over several ranks')
            self._batching = mode == 'autograd' and int(batch_size) < self.grid.shape[0]
        self._batch_generator = batch_generator
        _require_cuda(self.grid, 'grid')
        self.mode = mode
        self.weak_form = weak_form
        self.lambda_operator = lambda_operator
        self.lambda_bound = lambda_bound
        self.tol = tol
        self.derivative_points = derivative_points
        self.batch_size = int(batch_size) if self._batching else None
        # objects of the reference's own Equation_{NN,autograd,mat} classes (or anything with .operator / .bconds)
        # are re-wrapped: only the raw term dicts and conditions are read from them
        from .input_preprocessing import _EquationBase, Operator_bcond_preproc
        if not isinstance(equal_cls, _EquationBase):
            equal_cls = Operator_bcond_preproc(
                self.grid, equal_cls.operator, equal_cls.bconds, h=getattr(equal_cls, 'h', 0.001),
                inner_order=getattr(equal_cls, 'inner_order', '1'),
                boundary_order=getattr(equal_cls, 'boundary_order', '2')).set_strategy(mode)
        self.equal_cls = equal_cls
        self._shard = shard or (0, 1)
        self._pg = process_group
        self._nn_interior = nn_interior
        self._impl = impl
        equal_copy = deepcopy_equation(equal_cls)
        self._prepared_operator = equal_copy.operator_prepare()
        self.prepared_bconds = equal_copy.bnd_prepare()
        self._h = getattr(equal_cls, 'h', 0.001)
        self._inner_order = getattr(equal_cls, 'inner_order', '1')
        self._boundary_order = getattr(equal_cls, 'boundary_order', '2')
        self._fields_cache = None
        self._last_out = None
        self.loss = None
        self.loss_normalized = None
        self._build(model)
        from .eval import Operator, Bounds          # thin views over this object (compat shims)
        self.operator = Operator._from_solution(self)
        self.boundary = Bounds._from_solution(self)

    # -- plan construction -------------------------------------------------------------------------
    def _build(self, model):
        if self.mode == 'mat':
            from .mat import MatPlan
            self.model = check_device(model)
            _require_cuda(self.model, 'mat-mode model tensor')
            self._plan = MatPlan(self.grid, self._prepared_operator, self.prepared_bconds, self.model,
                                 self.lambda_operator, self.lambda_bound, self.derivative_points,
                                 shard=self._shard, process_group=self._pg, per_cell_coeffs=self.tol != 0)
            self._n_slots = self._plan.n_slots
            self.bval_keys = list(self._plan.bnd_types)
            self.bval_length = list(self._plan.type_len)
            return
        self.model = model.to(self.grid.device)
        for p in self.model.parameters():
            _require_cuda(p, 'model parameter')
        grid_rows = self.grid
        if self._batching:
            grid_rows = self.grid[:self.batch_size]          # the plan's interior segment holds one batch of rows
            for eq in self._prepared_operator:
                for term in eq.values():
                    c = term['coeff']
                    if isinstance(c, torch.Tensor) and not isinstance(c, torch.nn.Parameter) and c.numel() == self.grid.shape[0]:
                        raise UnsupportedProblem('per-point tensor coefficients with mini-batches (use a callable coefficient)')
        ir = lower_problem(self.mode, grid_rows, self._prepared_operator, self.prepared_bconds, self.model,
                           self.lambda_operator, self.lambda_bound, h=self._h, inner_order=self._inner_order,
                           boundary_order=self._boundary_order, nn_interior=self._nn_interior,
                           shard=self._shard)
        self._ir = ir
        if self._batching and ir.host_terms:
            raise UnsupportedProblem("callable 'pow' with mini-batches")
        self._plan = FusedPlan(ir, self.grid.device, impl=self._impl)
        if (self._shard[1] > 1 and self._collective == 'library' and self.tol == 0 and self.weak_form in (None, [])
                and _dist_ready()):
            # the all-reduce of [loss terms | gradient] moves under the C ABI (the causal loss keeps torch.distributed:
            # its forward-only launch must not be reduced)
            self._plan.comm_init(self._shard[0], self._shard[1], self._pg)
        self._n_slots = ir.n_slots
        self.bval_keys = list(ir.bnd_types)
        self.bval_length = list(ir.type_len)
        if self._batching:
            self._init_mini_batches()
            self.current_batch_i = 0

    # -- mini-batches (tedeous/eval.py:124-141) ----------------------------------------------------------------
    @property
    def n_batches(self) -> int:
        n, b = self.grid.shape[0], self.batch_size
        return (n + b - 1) // b if self._batching else 1

    def _init_mini_batches(self):
        """A fresh shuffle of the grid rows, consuming the generator exactly like the reference's
        iter(DataLoader(grid, batch_size, shuffle=True, generator=g)): one int64 (the iterator's base seed), then
        RandomSampler's torch.randperm(n, generator=g) - plus the sampler's empty tail randperm when an epoch ends - so the
        same generator state yields the same batches as the reference."""
        if self._batch_generator is None:
            self._batch_generator = torch.Generator(device=self.grid.device)
        g = self._batch_generator
        n = self.grid.shape[0]
        if getattr(self, '_batches', None) is not None:
            torch.randperm(n, generator=g, device=g.device)       # RandomSampler's tail draw when an epoch is exhausted
        torch.empty((), dtype=torch.int64, device=g.device).random_(generator=g)   # the DataLoader iterator's base seed
        perm = torch.randperm(n, generator=g, device=g.device).to(self.grid.device)
        self._batches = list(perm.split(self.batch_size))
        self._batch_next = 0

    def _next_batch(self):
        """Binds the next batch of rows to the plan (eval.py:174-182); after the last one the rows are reshuffled."""
        idx = self._batches[self._batch_next]
        self._batch_next += 1
        wrapped = self._batch_next == len(self._batches)
        if wrapped:
            self._init_mini_batches()
        n_valid = int(idx.numel())
        rows = self.grid[idx]
        if n_valid < self.batch_size:
            rows = torch.cat([rows, rows[-1:].expand(self.batch_size - n_valid, -1)])
        self._plan.set_interior_rows(rows, n_valid)
        return n_valid, wrapped

    def _model_change(self, new_model) -> None:
        """Swap the model (Cache callback, tedeous/solution.py:109-127): rebuilds the plan's parameter view."""
        self._build(new_model)
        self._fields_cache = None

    # -- evaluation --------------------------------------------------------------------------------
    def _run_plan(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """-> (out [2 + n_slots (+ n_params)], flat gradient)."""
        if self.mode == 'mat':
            if self.tol != 0:
                self._causal_weights_mat()
            return self._plan.loss_grad_raw(self.model)
        if self.tol != 0:
            self._causal_weights()
        out = self._plan.loss_grad()
        if self._shard[1] > 1 and not self._plan.has_comm:
            import torch.distributed as dist
            dist.all_reduce(out, op=dist.ReduceOp.SUM, group=self._pg)
        return out, out[2 + self._n_slots:]

    def capture_step(self):
        """CUDA graph of one whole loss + gradient step as `_run_plan` enqueues it: parameter packing, the interior and
        boundary launches with their fork / join, the reduction and - over several ranks - the NCCL all-reduce of the
        [loss terms | gradient] vector (mat mode: halo exchange, the kernel launches, all-reduce of the loss terms).
        -> (replay, out): `out` is a static tensor refreshed by every replay; parameter tensors must keep their storage
        (in-place optimiser updates do).  Steps that are launch bound (BASELINE config 1) replay as one graph launch."""
        if self.tol != 0 or self.weak_form not in (None, []) or self._hybrid:
            raise UnsupportedProblem("graph capture of the causal / weak-form / callable-'pow' step is not provided")
        if self.mode == 'mat':
            replay, out, self._graph_grad = self._plan.capture(self.model)
            return replay, out
        dev = self.grid.device
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(3):                          # warm-up outside the capture (function attributes, NCCL)
                self._run_plan()
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = self._run_plan()[0]
        self._graph = graph                             # keeps the graph (and its private pool) alive
        return graph.replay, out

    def _causal_weights(self):
        """Causal loss (tedeous/losses.py:137-182): w[t, j] = exp(-tol * sum_{s < t} res[s, j]) with
        res = sum_eq op^2 reshaped to [n_t, N / n_t] (column 0 of the grid is the slowest coordinate), not
        differentiated.  One forward-only launch yields op; the weights then enter the fused loss + gradient launch
        as per-row factors (`tdb200_plan_set_row_weights`); lambda_operator is not used by this loss."""
        self._plan.set_row_weights(None)
        fields, _ = self._plan.eval_fields()
        seg = self._ir.segments[0]
        n, ncols = seg.n_groups, len(seg.cols)
        op = fields[:n * ncols].reshape(n, ncols)
        w = causal_row_weights((op * op).sum(1), self.tol, self._ir.n_t, self._ir.n_interior,
                               getattr(seg, 'shard_range', None), self._shard[1], self._pg)
        self._plan.set_row_weights(w)

    def _causal_weights_mat(self):
        """Causal loss in mat mode (losses.py:137-182 with n_t = grid.shape[1], solution.py:58-60): the no-grad weights
        w[t, j] = exp(-tol * sum_{s < t} sum_eq op^2[s, j]) over the grid rows of axis 0, from a forward-only launch with the
        plain coefficients; they enter the loss + gradient launch as a per-cell factor sqrt(w) of every equation
        coefficient (`MatPlan.set_cell_weights`); lambda_operator is not used by this loss."""
        plan = self._plan
        plan.set_cell_weights(None)
        op = plan.eval_fields(self.model)[0]
        n0, n1 = plan.ir.shape_ext[1], plan.ir.shape_ext[2]
        res = (op[:n0 * n1] ** 2).sum(1).reshape(n0, n1)
        w = torch.exp(-self.tol * (torch.cumsum(res, 0) - res))
        plan.set_cell_weights(w)

    def _sync_lambdas(self):
        """Push lambda_operator / lambda_bound to the plan when they changed (callbacks such as AdaptiveLambda assign new
        tensors).  Reading a device tensor back is a host sync, so unchanged objects (same identity and version counter)
        are not looked at again: the steady-state step has no sync before the launch."""
        objs = (self.lambda_operator, self.lambda_bound)
        vers = tuple(getattr(x, '_version', None) for x in objs)
        seen = getattr(self, '_lam_seen', None)               # holds references: an id cannot be recycled while cached
        if seen is not None and all(a is b for a, b in zip(seen[0], objs)) and seen[1] == vers:
            return
        self._lam_seen = (objs, vers)
        n_eq = self._n_slots - len(self.bval_keys)

        def as_list(lam, n):
            if isinstance(lam, torch.Tensor):
                return [float(x) for x in lam.detach().reshape(-1).cpu().tolist()]
            if isinstance(lam, (int, float)):
                return [float(lam)] * n
            return [float(x) for x in lam]
        lam_op = as_list(1 if self.tol != 0 else self.lambda_operator, n_eq)
        lam_b = as_list(self.lambda_bound, len(self.bval_keys))
        lam = lam_op + lam_b
        if len(lam) != self._n_slots:
            raise ValueError(f'expected {self._n_slots} lambdas, got {len(lam)}')
        if lam != list(self._plan.slot_lambda):
            self._plan.set_lambdas(lam)

    def _evaluate_weak(self, save_graph: bool) -> Tuple[torch.Tensor, torch.Tensor]:
        """Weak-form loss (tedeous/losses.py:184-228 over eval.py:195-221): per-point fields from a forward launch,
        the nested integrals as a few device ops, the parameter gradient by the fused kernel in seed mode."""
        from .losses import Losses, weak_operator
        self._fields_cache = None
        grad = save_graph and torch.is_grad_enabled()
        op, bval, tval = self._fields_differentiable() if grad else self._fields()
        self._fields_cache = (op.detach(), bval.detach(), tval)
        wop = weak_operator(op, self._ir.interior_points, self.weak_form)
        n_eq = self._n_slots - len(self.bval_keys)
        loss, loss_n = Losses(self.mode, self.weak_form, None, 0).compute(
            wop, bval, tval, self.lambda_operator, self.lambda_bound, save_graph)
        self.loss, self.loss_normalized = loss, loss_n.detach()
        self.lambda_operator = lambda_prepare(torch.empty(1, n_eq), self.lambda_operator).to(torch.float32)
        self.lambda_bound = lambda_prepare(torch.empty(1, len(self.bval_keys)), self.lambda_bound).to(torch.float32)
        return self.loss, self.loss_normalized

    # -- callable 'pow' (tedeous/derivative.py:52-55, 126-129) ---------------------------------------------------
    @property
    def _hybrid(self) -> bool:
        return self.mode != 'mat' and bool(getattr(self._ir, 'host_terms', None))

    def _assemble_host(self, op_ext: torch.Tensor) -> torch.Tensor:
        """[N, n_eq + factor fields] from the kernel -> op [N, n_eq]: the callable-'pow' terms are evaluated by torch on
        the factor columns exactly as the reference chains them (`der_term = pow_j(der_term * factor_j)` for a callable,
        `der_term * factor_j ** pow_j` for a number) and added to their equation's column."""
        n_eq = self._ir.n_eq
        cols = [op_ext[:, e] for e in range(n_eq)]
        for e, coeff, chain in self._ir.host_terms:
            der = 1.
            for col, pw in chain:
                val = op_ext[:, col]
                der = pw(der * val) if callable(pw) else der * val ** pw
            c = coeff.to(op_ext.device).reshape(-1) if isinstance(coeff, torch.Tensor) and coeff.numel() > 1 else coeff
            cols[e] = cols[e] + c * der
        return torch.stack(cols, 1)

    def _evaluate_hybrid(self, save_graph: bool) -> Tuple[torch.Tensor, torch.Tensor]:
        """Default / causal loss of a problem with callable powers: per-point fields from a forward launch, the callables
        and the loss formula as device ops under autograd, the parameter gradient by the fused kernel in seed mode."""
        from .losses import Losses
        self._fields_cache = None
        grad = save_graph and torch.is_grad_enabled()
        op, bval, tval = self._fields_differentiable() if grad else self._fields()
        self._fields_cache = (op.detach(), bval.detach(), tval)
        n_eq = self._n_slots - len(self.bval_keys)
        n_t = None
        if self.tol != 0:
            n_t = int(torch.unique(self._ir.interior_points[:, 0]).numel())
        loss, loss_n = Losses(self.mode, None, n_t, self.tol).compute(
            op, bval, tval, self.lambda_operator, self.lambda_bound, save_graph)
        self.loss, self.loss_normalized = loss.reshape(1), loss_n.detach().reshape(1)
        self._last_out = torch.cat([self.loss.detach(), self.loss_normalized,
                                    torch.mean(op.detach() ** 2, 0), torch.mean((bval.detach() - tval) ** 2, 0)])
        self.lambda_operator = lambda_prepare(torch.empty(1, n_eq), self.lambda_operator).to(torch.float32)
        self.lambda_bound = lambda_prepare(torch.empty(1, len(self.bval_keys)), self.lambda_bound).to(torch.float32)
        return self.loss, self.loss_normalized

    def refresh_coeffs(self) -> None:
        """Re-evaluate callable coefficients (modes NN / autograd) into the plan's buffers."""
        if self.mode != 'mat':
            self._plan.refresh_coeffs()

    def evaluate(self, save_graph: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
        """One loss evaluation (tedeous/solution.py:129-168)."""
        if self._callable_coeffs == 'every_step':
            self.refresh_coeffs()
        if self.weak_form not in (None, []):
            return self._evaluate_weak(save_graph)
        if self._hybrid:
            return self._evaluate_hybrid(save_graph)
        self._sync_lambdas()
        self._fields_cache = None
        if self._batching:
            n_valid, wrapped = self._next_batch()
        if self.mode == 'mat':
            params = [self.model]
        else:
            params = self._plan.ir.net.param_tensors()
        needs_grad = save_graph and torch.is_grad_enabled() and any(p.requires_grad for p in params)
        if needs_grad:
            loss = _FusedLoss.apply(self, *params)
        else:
            self._last_out, self._last_grad = self._run_plan()
            loss = self._last_out[0:1].clone()
        self.loss = loss
        self.loss_normalized = self._last_out[1:2].clone()
        n_eq = self._n_slots - len(self.bval_keys)
        dt = torch.float32
        if not isinstance(self.lambda_operator, torch.Tensor) or self.lambda_operator.dtype != dt:
            self.lambda_operator = lambda_prepare(torch.empty(1, n_eq), self.lambda_operator).to(dt)
        if not isinstance(self.lambda_bound, torch.Tensor) or self.lambda_bound.dtype != dt:
            self.lambda_bound = lambda_prepare(torch.empty(1, len(self.bval_keys)), self.lambda_bound).to(dt)
        if self._batching:
            # solution.py:159-166: the residual fields of the epoch's batches are concatenated in `save_op`
            op = self._fields()[0][:n_valid].detach()
            self.save_op = op if self.current_batch_i == 0 else torch.cat((self.save_op, op), 0)
            self.current_batch_i = 0 if wrapped else self.current_batch_i + 1
            self.operator.current_batch_i = self.current_batch_i
        return self.loss, self.loss_normalized

    # per-column mean squares of the last evaluation (the natural partial results)
    @property
    def op_mse(self) -> torch.Tensor:
        n_eq = self._n_slots - len(self.bval_keys)
        return self._last_out[2:2 + n_eq]

    @property
    def bval_mse(self) -> torch.Tensor:
        n_eq = self._n_slots - len(self.bval_keys)
        return self._last_out[2 + n_eq:2 + self._n_slots]


    # -- per-point fields on demand ----------------------------------------------------------------
    def _fields(self):
        if self._fields_cache is None:
            if self._shard[1] > 1:
                raise UnsupportedProblem('per-point fields are not gathered across ranks')
            if self.mode == 'mat':
                if self.tol != 0:
                    self._plan.set_cell_weights(None)          # fields are those of the plain coefficients
                self._fields_cache = self._plan.eval_fields(self.model)
            else:
                op, bval, tval = _assemble_fields(self._plan)
                self._fields_cache = (self._assemble_host(op) if self._hybrid else op, bval, tval)
        return self._fields_cache

    def _fields_differentiable(self):
        """(op, bval, true_bval) with op / bval attached to the autograd graph of the network parameters."""
        if self.mode == 'mat':
            raise UnsupportedProblem('differentiable per-point fields are implemented for modes NN / autograd')
        if self._shard[1] > 1:
            raise UnsupportedProblem('per-point fields are not gathered across ranks')
        params = self._plan.ir.net.param_tensors()
        op, bval, tval = _assemble_fields(self._plan, _FusedFields.apply(self._plan, *params))
        return (self._assemble_host(op) if self._hybrid else op), bval, tval

    def residual_jacobian(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """(J_op [N * n_eq, P], J_bnd [max_len * n_types, P]): the Jacobians of `op.reshape(-1)` and of
        `(bval - true_bval).reshape(-1)` with respect to the flat parameter vector (`parameters_to_vector(model.parameters())`
        order - trainable coefficients registered on the net come first there; rows
        of the zero padding of `bval` are zero) - what `NGD.gram_factory` (tedeous/optimizers/ngd.py:57-77) assembles
        from one `autograd.grad` per residual.  One SIMT launch per (segment, residual column)."""
        if self.mode == 'mat':
            raise UnsupportedProblem('per-residual Jacobian rows are implemented for modes NN / autograd')
        if self._shard[1] > 1:
            raise UnsupportedProblem('per-residual Jacobian rows are not gathered across ranks')
        if getattr(self, '_batching', False):
            raise UnsupportedProblem('per-residual Jacobian rows with mini-batches')
        if self._hybrid:
            raise UnsupportedProblem("per-residual Jacobian rows with callable 'pow' terms")
        plan, ir = self._plan, self._plan.ir
        n_types, max_len, P = len(ir.bnd_types), max(ir.type_len), plan.n_params
        j_op = None
        j_bnd = torch.zeros(max_len, n_types, P, dtype=torch.float32, device=plan.device)
        for si, s in enumerate(ir.segments):
            if s.slots[0] < ir.n_eq:
                cols = [plan.jacobian_rows(si, c) for c in range(len(s.cols))]
                j_op = torch.stack(cols, dim=1).reshape(-1, P)               # row-major [N, n_eq] like op.reshape(-1)
            else:
                idx = s.row_index.to(plan.device)
                j_bnd[idx, s.slots[0] - ir.n_eq] = plan.jacobian_rows(si, 0)
        perm = self._param_order()
        j_bnd = j_bnd.reshape(-1, P)
        return (j_op, j_bnd) if perm is None else (j_op[:, perm], j_bnd[:, perm])

    def _param_order(self):
        """Columns of the plan's flat gradient (W0, b0, ..., trainable coefficients last) in the order of
        `model.parameters()` / `parameters_to_vector` (parameters registered on the net itself come FIRST there,
        tedeous/models.py:183-195); None if the orders agree."""
        offs, o = {}, 0
        for p in self._plan.ir.net.param_tensors():
            offs[id(p)] = (o, p.numel())
            o += p.numel()
        idx = [torch.arange(*(lambda sn: (sn[0], sn[0] + sn[1]))(offs[id(p)]), device=self._plan.device)
               for p in self.model.parameters()]
        idx = torch.cat(idx)
        return None if torch.equal(idx, torch.arange(idx.numel(), device=idx.device)) else idx

    def residual_jvp(self, v: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """(J_op v, J_bnd v) for a flat parameter-space vector v: the directional derivative of every residual."""
        j_op, j_bnd = self.residual_jacobian()
        v = v.detach().to(j_op.device, torch.float32).reshape(-1)
        return j_op @ v, j_bnd @ v

    @property
    def op(self) -> torch.Tensor:
        return self._fields()[0]

    @property
    def bval(self) -> torch.Tensor:
        return self._fields()[1]

    @property
    def true_bval(self) -> torch.Tensor:
        return self._fields()[2]


def _assemble_fields(plan: FusedPlan, fields: Optional[torch.Tensor] = None):
    """op [N, n_eq], bval / true_bval [max_len, n_types] zero padded (eval.py:55-87).  `fields`: an already evaluated
    (possibly differentiable) flat field vector; default: one forward launch."""
    ir = plan.ir
    if fields is None:
        fields, _ = plan.eval_fields()
    seg_rec = plan.flat.seg
    n_types = len(ir.bnd_types)
    max_len = max(ir.type_len)
    bval = torch.zeros(max_len, n_types, dtype=torch.float32, device=plan.device)
    tval = torch.zeros_like(bval)
    op = None
    for s, rec in zip(ir.segments, seg_rec):
        off, n, nc = int(rec['field_off']), s.n_groups, len(s.cols)
        vals = fields[off:off + n * nc].reshape(n, nc)
        if s.slots[0] < ir.n_eq:
            op = vals
        else:
            col = s.slots[0] - ir.n_eq
            idx = s.row_index.to(plan.device)
            bval = bval.index_put((idx, torch.full_like(idx, col)), vals[:, 0])      # out of place: keeps the graph
            tval[idx, col] = s.targets.reshape(-1).to(device=plan.device, dtype=torch.float32)
    return op, bval, tval


def deepcopy_equation(equal_cls):
    """deepcopy that keeps tensors / nn.Parameters / callables by reference (tedeous/solution.py:62, 90-107:
    trainable coefficients must stay the live Parameter objects)."""
    memo = {}

    def keep(obj):
        memo[id(obj)] = obj

    def walk(o):
        if isinstance(o, torch.Tensor) or callable(o) and not isinstance(o, type):
            keep(o)
        elif isinstance(o, dict):
            for v in o.values():
                walk(v)
        elif isinstance(o, (list, tuple)):
            for v in o:
                walk(v)
    walk(getattr(equal_cls, 'operator', None))
    walk(getattr(equal_cls, 'bconds', None))
    keep(equal_cls.grid)
    return deepcopy(equal_cls, memo)
