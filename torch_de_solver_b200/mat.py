"""mat mode: the solution is a tensor on a tensor-product grid and derivatives are finite-difference
stencils (tedeous/derivative.py:135-323, eval.py:143-193 / 283-461 mat branches, input_preprocessing.py:511-594).

Lowering done here, once:
* every factor `D_a^k u_v` becomes a *derivative field* with a banded 1-D matrix.  The reference's first
  derivative is the average of a backward and a forward `p`-point rule in the interior and the one-sided rule
  on the first / last `p - 1` nodes (derivative.py:199-291, coefficients from a Vandermonde solve 174-197);
  higher orders are repeated applications, so `D^k` has half-width `k (p - 1)` and `k (p - 1)` special rows at
  each end.  The band is built by composing the dense first-derivative matrix in fp64
  and dividing by the fp32 grid step `h` the reference computes (derivative.py:229-247);
* boundary points become flat cell indices by index arithmetic on the axis coordinates (the reference scans
  the whole grid with isclose per boundary point - input_preprocessing.py:553-575).
The fused kernel (csrc/mat_stencil.cu) then evaluates residual, loss and d loss / d u in one pass."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _native
from .plan import FACTOR_DTYPE, TERM_DTYPE, UnsupportedProblem, _lambda_list

FIELD_DTYPE = np.dtype([('var', '<i4'), ('axis', '<i4'), ('order', '<i4'), ('half_width', '<i4'),
                        ('n_edge', '<i4'), ('coef_off', '<i4')], align=True)


def first_derivative_matrix(n: int, p: int) -> np.ndarray:
    """Dense D (h = 1) exactly as Derivative_mat._derivative builds it along one axis."""
    back = list(range(-p + 1, 1))
    farw = list(range(p))

    def alpha(labels):
        lab = np.array(labels, dtype=np.float64)
        A = np.array([lab ** i for i in range(len(labels))])
        b = np.zeros(len(labels))
        b[1] = 1
        return np.linalg.solve(A, b)
    ab, af = alpha(back), alpha(farw)
    D = np.zeros((n, n))
    for r in range(n):
        if r < p - 1:
            for lab, a in zip(farw, af):
                D[r, r + lab] += a
        elif r >= n - (p - 1):
            for lab, a in zip(back, ab):
                D[r, r + lab] += a
        else:
            for lab, a in zip(back, ab):
                D[r, r + lab] += a / 2
            for lab, a in zip(farw, af):
                D[r, r + lab] += a / 2
    return D


def derivative_band(n: int, p: int, order: int, h: float) -> Tuple[np.ndarray, int, int]:
    """-> (band floats: interior[2b+1], lo[E][2b+1], hi[E][2b+1]; b; E) for D^order / h^order on n nodes."""
    b = order * (p - 1)
    if n < 2 * p:
        raise UnsupportedProblem(f'axis with {n} nodes is too short for {p}-point derivatives')
    E = b
    n_small = min(n, 4 * b + 2 * p + 3)
    if n <= 2 * E + 2 * b + 1:           # short axis: every row gets its own coefficients
        n_small = n
        E = (n + 1) // 2
    Dk = np.linalg.matrix_power(first_derivative_matrix(n_small, p), order) / (h ** order)
    w = 2 * b + 1

    def row_band(r):
        out = np.zeros(w)
        for m in range(-b, b + 1):
            if 0 <= r + m < n_small:
                out[m + b] = Dk[r, r + m]
        # nothing outside the band
        mask = np.ones(n_small, bool)
        mask[max(0, r - b):min(n_small, r + b + 1)] = False
        assert np.all(np.abs(Dk[r][mask]) < 1e-9 * (np.abs(Dk[r]).max() + 1e-300)), 'band too narrow'
        return out
    interior = row_band(n_small // 2) if n_small > 2 * E else np.zeros(w)
    if n_small > 2 * E:          # interior rows must be uniform
        for r in range(E, n_small - E):
            assert np.allclose(row_band(r), interior, rtol=1e-12, atol=1e-12 * np.abs(interior).max())
    lo = np.stack([row_band(r) for r in range(E)])
    hi = np.stack([row_band(n_small - 1 - r) for r in range(E)])
    return np.concatenate([interior, lo.reshape(-1), hi.reshape(-1)]).astype(np.float32), b, E


def _axis_coords(grid: torch.Tensor) -> List[torch.Tensor]:
    d = grid.shape[0]
    out = []
    for a in range(d):
        idx = [0] * d
        idx[a] = slice(None)
        out.append(grid[(a, *idx)].contiguous())
    return out


def step_h(grid: torch.Tensor) -> List[float]:
    """|unique(coord)[1] - unique(coord)[0]| in fp32 (derivative.py:229-247) without the full-grid unique."""
    hs = []
    for c in _axis_coords(grid):
        u = torch.unique(c.float())
        hs.append(float(abs(u[1] - u[0])))
    return hs


def cell_indices(grid: torch.Tensor, bnd: torch.Tensor) -> torch.Tensor:
    """Flat cell index of every boundary point (nearest node per axis, checked with isclose like
    input_preprocessing.py:553-575)."""
    axes = _axis_coords(grid)
    shape = grid.shape[1:]
    flat = torch.zeros(bnd.shape[0], dtype=torch.int64, device=bnd.device)
    for a, coords in enumerate(axes):
        c = coords.float()
        x = bnd[:, a].float()
        order = torch.argsort(c)
        cs = c[order]
        pos = torch.searchsorted(cs, x).clamp(1, cs.numel() - 1)
        left = (x - cs[pos - 1]).abs() <= (cs[pos] - x).abs()
        near = torch.where(left, pos - 1, pos)
        if not bool(torch.isclose(cs[near], x).all()):
            raise ValueError('a boundary point does not lie on the grid')
        flat = flat * shape[a] + order[near]
    return flat.to(torch.int32)


class MatPlan:
    """Owns one tdb200_mat_plan."""

    def __init__(self, grid: torch.Tensor, prepared_operator: List[dict], bconds: List[dict], model: torch.Tensor,
                 lambda_operator, lambda_bound, derivative_points: int = 2, shard=(0, 1), process_group=None):
        if shard[1] > 1:
            raise UnsupportedProblem('multi-GPU slab decomposition of mat mode is not implemented yet')
        if grid.dim() != 3 or model.dim() != 3:
            raise UnsupportedProblem('the fused mat path supports 2-D grids ([2, N0, N1], model [n_eq, N0, N1])')
        if model.dtype != torch.float32:
            raise UnsupportedProblem('mat-mode model must be float32')
        if not bconds:
            raise UnsupportedProblem('a problem without boundary conditions has no finite loss in the reference')
        self.lib = _native.load()
        self.device = model.device
        self.grid = grid
        n_var, n0, n1 = model.shape
        self.shape = (n_var, n0, n1)
        n_eq = len(prepared_operator)
        p = derivative_points
        hs = step_h(grid)
        dims = (n0, n1)

        self._fields: List[Tuple[int, int, int]] = [(v, 0, 0) for v in range(n_var)]
        terms, factors, coefs = [], [], []
        coef_off = 0

        def field_index(var, axes):
            axes = [a for a in axes if a is not None]
            if not axes:
                return var
            if len(set(axes)) != 1:
                raise UnsupportedProblem(f'mixed partial derivative {axes} is not supported by the fused mat path')
            key = (var, axes[0], len(axes))
            if key not in self._fields:
                self._fields.append(key)
            return self._fields.index(key)

        def add_terms(op: dict):
            nonlocal coef_off
            begin = len(terms)
            for label, term in op.items():
                dif = list(term.keys())[1]
                fb = len(factors)
                for spec, pw, var in zip(term[dif], term['pow'], term['var']):
                    if callable(pw):
                        raise UnsupportedProblem("callable 'pow' is not supported by the fused path")
                    spec = [] if spec == [None] else spec
                    ip = int(pw) if float(pw).is_integer() and 0 <= pw <= 16 else -1
                    factors.append((int(var), field_index(int(var), spec), float(pw), ip))
                    if var >= n_var:
                        raise ValueError(f'var {var} but the model has {n_var} fields')
                c = term['coeff']
                if isinstance(c, torch.nn.Parameter):
                    raise UnsupportedProblem('trainable coefficients in mat mode')
                if callable(c) and not isinstance(c, torch.Tensor):
                    c = c(grid)
                if isinstance(c, torch.Tensor) and c.numel() > 1:
                    c = torch.broadcast_to(c.to(self.device, torch.float32), (n0, n1)).reshape(-1)
                    terms.append((0.0, 1, coef_off, fb, len(factors)))
                    coefs.append(c)
                    coef_off += c.numel()
                else:
                    terms.append((float(c), 0, 0, fb, len(factors)))
            return begin, len(terms)

        eq_ranges = [add_terms(eq) for eq in prepared_operator]

        # ---- boundary rows ----------------------------------------------------------------------------
        self.bnd_types: List[str] = []
        type_len: Dict[str, int] = {}
        bc_rows, cells, targets = [], [], []
        cell_off = tgt_off = 0
        self._bc_layout = []                       # (type index, offset in type column, n) per condition
        for bc in bconds:
            kind = bc['type']
            if kind not in self.bnd_types:
                self.bnd_types.append(kind)
                type_len[kind] = 0
            slot = self.bnd_types.index(kind)
            bop = bc['bop']
            if kind == 'robin':
                raise UnsupportedProblem('robin conditions in mat mode')
            if kind == 'periodic':
                sides = [cell_indices(grid, b) for b in bc['bnd']]
                K = len(sides)
                if K > 4:
                    raise UnsupportedProblem('periodic condition with more than 4 sides')
                n = sides[0].numel()
                cidx = torch.stack(sides, 1).reshape(-1)
                sign = [1.0] + [-1.0] * (K - 1) + [0.0] * (4 - K)
                tgt = torch.zeros(n, dtype=torch.float32, device=self.device)
            else:
                cidx = cell_indices(grid, bc['bnd'])
                n, K = cidx.numel(), 1
                sign = [1.0, 0.0, 0.0, 0.0]
                tgt = bc['bval'].reshape(-1).to(self.device, torch.float32)
                if tgt.numel() != n:
                    raise ValueError(f'{tgt.numel()} target values for {n} boundary points')
            tb = te = 0
            if bop is not None:
                tb, te = add_terms(bop)
            bc_rows.append((n, cell_off, tgt_off, K, int(bc['var']), slot, tb, te, sign))
            self._bc_layout.append((slot, type_len[kind], n))
            cells.append(cidx)
            targets.append(tgt)
            cell_off += n * K
            tgt_off += n
            type_len[kind] += n
        self.type_len = [type_len[t] for t in self.bnd_types]
        max_len = max(self.type_len)
        self.n_eq = n_eq
        self.n_slots = n_eq + len(self.bnd_types)
        self.slot_len = [n0 * n1] * n_eq + [max_len] * len(self.bnd_types)
        self.slot_lambda = _lambda_list(lambda_operator, n_eq, 'lambda_operator') + \
            _lambda_list(lambda_bound, len(self.bnd_types), 'lambda_bound')

        # ---- bands for every derivative field ------------------------------------------------------------
        if len(self._fields) > 12:
            raise UnsupportedProblem('more than 12 distinct derivative fields')
        fld = np.zeros(len(self._fields), FIELD_DTYPE)
        band = [np.zeros(1, np.float32)]
        off = 1
        for q, (var, axis, order) in enumerate(self._fields):
            fld[q]['var'], fld[q]['axis'], fld[q]['order'] = var, axis, order
            if order > 0:
                bnd_arr, b, E = derivative_band(dims[axis], p, order, hs[axis])
                fld[q]['half_width'], fld[q]['n_edge'], fld[q]['coef_off'] = b, E, off
                band.append(bnd_arr)
                off += bnd_arr.size
        band = np.concatenate(band).astype(np.float32)

        self._terms = np.array(terms, dtype=TERM_DTYPE) if terms else np.zeros(0, TERM_DTYPE)
        self._factors = np.array(factors, dtype=FACTOR_DTYPE) if factors else np.zeros(0, FACTOR_DTYPE)
        desc = _native.MatDesc(n_eq, n_var, n0, n1, len(self._fields))
        eb = np.array([r[0] for r in eq_ranges], np.int32)
        ee = np.array([r[1] for r in eq_ranges], np.int32)
        handle = C.c_void_p()
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        _native.check(self.lib.tdb200_mat_plan_create(
            C.byref(desc), _native.np_ptr(fld), band.size, _native.np_ptr(band), _native.np_ptr(eb), _native.np_ptr(ee),
            len(self._terms), _native.np_ptr(self._terms), len(self._factors), _native.np_ptr(self._factors),
            dev_index, C.byref(handle)), 'tdb200_mat_plan_create')
        self.handle = handle
        self._coeffs = torch.cat(coefs).contiguous() if coefs else torch.zeros(1, device=self.device)
        _native.check(self.lib.tdb200_mat_plan_set_coeffs(handle, self._coeffs.data_ptr(), self._coeffs.numel()),
                      'tdb200_mat_plan_set_coeffs')
        self._bcs = np.zeros(len(bc_rows), _native.MAT_BC_DTYPE)
        for i, (n, co, to, K, var, slot, tb, te, sign) in enumerate(bc_rows):
            r = self._bcs[i]
            r['n_rows'], r['cell_off'], r['tgt_off'], r['K'], r['var'], r['slot'] = n, co, to, K, var, slot
            r['term_begin'], r['term_end'] = tb, te
            r['sign'][:] = sign
        self._cells = torch.cat(cells).to(self.device, torch.int32).contiguous()
        self._targets = torch.cat(targets).contiguous()
        self.n_bc_rows = int(self._targets.numel())
        self._push_bcs()
        self.out_size = int(self.lib.tdb200_mat_plan_out_size(handle))
        self.launches_per_call = int(self.lib.tdb200_mat_plan_launches_per_call(handle))
        self.kernel_kind = ('generic', 'register-tap', 'cross-vec4', 'cross-tma')[int(self.lib.tdb200_mat_plan_kernel_kind(handle))]
        self.n_cells = n0 * n1

    def _push_bcs(self):
        lam = np.asarray(self.slot_lambda, np.float64)
        ln = np.asarray(self.slot_len, np.float64)
        _native.check(self.lib.tdb200_mat_plan_set_bcs(
            self.handle, len(self._bcs), _native.np_ptr(self._bcs), self._cells.data_ptr(), self._targets.data_ptr(),
            self.n_slots, _native.np_ptr(lam), _native.np_ptr(ln)), 'tdb200_mat_plan_set_bcs')

    def set_lambdas(self, slot_lambda):
        self.slot_lambda = [float(x) for x in slot_lambda]
        self._push_bcs()

    def _check_model(self, u):
        if tuple(u.shape) != self.shape or u.dtype != torch.float32 or not u.is_cuda or not u.is_contiguous():
            raise RuntimeError(f'mat-mode model must be a contiguous float32 CUDA tensor of shape {self.shape}')

    def loss_grad_raw(self, u: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        self._check_model(u)
        out = torch.empty(self.out_size, dtype=torch.float32, device=self.device)
        grad = torch.empty(self.shape, dtype=torch.float32, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _native.check(self.lib.tdb200_mat_loss_grad(self.handle, u.data_ptr(), grad.data_ptr(), out.data_ptr(),
                                                    stream), 'tdb200_mat_loss_grad')
        return out, grad

    def eval_fields(self, u: torch.Tensor):
        self._check_model(u)
        out = torch.empty(self.out_size, dtype=torch.float32, device=self.device)
        op = torch.empty(self.n_cells, self.n_eq, dtype=torch.float32, device=self.device)
        rows = torch.empty(max(self.n_bc_rows, 1), dtype=torch.float32, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _native.check(self.lib.tdb200_mat_eval_fields(self.handle, u.data_ptr(), op.data_ptr(), rows.data_ptr(),
                                                      out.data_ptr(), stream), 'tdb200_mat_eval_fields')
        max_len = max(self.type_len)
        bval = torch.zeros(max_len, len(self.bnd_types), dtype=torch.float32, device=self.device)
        tval = torch.zeros_like(bval)
        off = 0
        for slot, base, n in self._bc_layout:
            bval[base:base + n, slot] = rows[off:off + n]
            tval[base:base + n, slot] = self._targets[off:off + n]
            off += n
        return op, bval, tval

    def __del__(self):
        try:
            if getattr(self, 'handle', None):
                self.lib.tdb200_mat_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass
