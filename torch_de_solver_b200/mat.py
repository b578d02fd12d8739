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
        hs.append(float(abs(u[1] - u[0])) if u.numel() > 1 else 1.0)     # (dummy axis of a lifted 1-D grid: unused)
    return hs


def cell_indices(grid: torch.Tensor, bnd: torch.Tensor) -> torch.Tensor:
    """Flat cell index of every boundary point (nearest node per axis, checked with isclose like
    input_preprocessing.py:553-575)."""
    axes = _axis_coords(grid)
    shape = grid.shape[1:]
    flat = torch.zeros(bnd.shape[0], dtype=torch.int64, device=bnd.device)
    for a, coords in enumerate(axes):
        c = coords.float()
        if c.numel() == 1 or a >= bnd.shape[1]:            # dummy axis of a lifted 1-D grid
            flat = flat * shape[a]
            continue
        x = bnd[:, a].float().contiguous()
        order = torch.argsort(c)
        cs = c[order]
        pos = torch.searchsorted(cs, x).clamp(1, cs.numel() - 1)
        left = (x - cs[pos - 1]).abs() <= (cs[pos] - x).abs()
        near = torch.where(left, pos - 1, pos)
        if not bool(torch.isclose(cs[near], x).all()):
            raise ValueError('a boundary point does not lie on the grid')
        flat = flat * shape[a] + order[near]
    return flat.to(torch.int32)


def slab_rows(n0: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of grid rows (axis 0) owned by `rank`."""
    return (n0 * rank) // world, (n0 * (rank + 1)) // world


class MatIR:
    """Lowered mat-mode problem of one rank (host data only; `MatPlan` hands it to the C ABI, tests/mat_interp.py
    evaluates it with dense fp64 operators).  With `shard = (rank, world)` the rank owns rows [r0, r1) of axis 0 and
    works on the extended slab [e0, e1) = owned rows + 2 * reach halo rows towards every neighbour: residuals of the
    halo rows next to the owned block are recomputed locally, so the only exchange per step is the 2 * reach rows of
    `u` from each neighbour plus one all-reduce of the loss terms."""

    def __init__(self, grid: torch.Tensor, prepared_operator: List[dict], bconds: List[dict], n_var: int,
                 device, lambda_operator, lambda_bound, derivative_points: int = 2, shard=(0, 1),
                 per_cell_coeffs: bool = False):
        self.lifted = grid.dim() == 2                      # 1-D grid [1, N0] -> [2, N0, 1] with a dummy second axis
        grid_user = grid                                   # what callable coefficients are evaluated on (derivative.py:319)
        if self.lifted:
            x = grid[0].reshape(-1, 1)
            grid = torch.stack([x, torch.zeros_like(x)])
        if grid.dim() != 3:
            raise UnsupportedProblem('the fused mat path supports 1-D and 2-D grids ([d, N0(, N1)], model [n_eq, N0(, N1)])')
        if not bconds:
            raise UnsupportedProblem('a problem without boundary conditions has no finite loss in the reference')
        rank, world = shard
        self.shard = (int(rank), int(world))
        self.device = device
        n0, n1 = int(grid.shape[1]), int(grid.shape[2])
        self.n0_global, self.n1, self.n_var = n0, n1, n_var
        n_eq = len(prepared_operator)
        p = derivative_points
        hs = step_h(grid)

        self.fields: List[Tuple[int, int, int]] = [(v, 0, 0) for v in range(n_var)]
        terms, factors, coefs_full = [], [], []

        def field_index(var, axes):
            axes = [a for a in axes if a is not None]
            if not axes:
                return var
            if len(set(axes)) != 1:
                raise UnsupportedProblem(f'mixed partial derivative {axes} is not supported by the fused mat path')
            key = (var, axes[0], len(axes))
            if key not in self.fields:
                self.fields.append(key)
            return self.fields.index(key)

        def add_terms(op: dict, per_cell: bool = False):
            begin = len(terms)
            for label, term in op.items():
                dif = list(term.keys())[1]
                fb = len(factors)
                for spec, pw, var in zip(term[dif], term['pow'], term['var']):
                    if isinstance(var, (list, tuple)) and len(var) == 1:
                        var = var[0]         # 'var': [0] next to a scalar 'pow' (equation_unify wraps it once more)
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
                    c = c(grid_user)
                if per_cell and not (isinstance(c, torch.Tensor) and c.numel() > 1):
                    c = torch.full((n0, n1), float(c), dtype=torch.float32, device=device)   # a buffer for every term
                if isinstance(c, torch.Tensor) and c.numel() > 1:
                    terms.append([0.0, 1, len(coefs_full), fb, len(factors)])       # idx = buffer number for now
                    c = c.to(device, torch.float32)
                    if self.lifted:                                                 # [1, N0] / [N0] values of a 1-D grid
                        c = c.reshape(-1, 1) if c.numel() == n0 else c
                    coefs_full.append(torch.broadcast_to(c, (n0, n1)))
                else:
                    terms.append([float(c), 0, 0, fb, len(factors)])
            return begin, len(terms)

        # per_cell_coeffs (causal loss): every equation term reads its coefficient from a per-cell buffer, so that the
        # no-grad causal weights can enter as a per-cell factor of the coefficients (MatPlan.set_cell_weights)
        self.eq_ranges = [add_terms(eq, per_cell_coeffs) for eq in prepared_operator]
        self.n_eq_buffers = len(coefs_full) if per_cell_coeffs else 0     # the first buffers belong to the equations

        # ---- boundary rows (global), operator terms of the conditions --------------------------------
        self.bnd_types: List[str] = []
        type_len: Dict[str, int] = {}
        bc_global = []
        for bc in bconds:
            kind = bc['type']
            if kind not in self.bnd_types:
                self.bnd_types.append(kind)
                type_len[kind] = 0
            slot = self.bnd_types.index(kind)
            bop = bc['bop']
            if kind == 'robin':
                bop = _robin_operator(bop, int(bc['var']))
            if kind == 'periodic':
                sides = [cell_indices(grid, b) for b in bc['bnd']]
                K = len(sides)
                if K > 4:
                    raise UnsupportedProblem('periodic condition with more than 4 sides')
                n = sides[0].numel()
                cidx = torch.stack(sides, 1)                        # [n, K]
                sign = [1.0] + [-1.0] * (K - 1) + [0.0] * (4 - K)
                tgt = torch.zeros(n, dtype=torch.float32, device=device)
            else:
                cidx = cell_indices(grid, bc['bnd']).reshape(-1, 1)
                n, K = cidx.shape[0], 1
                sign = [1.0, 0.0, 0.0, 0.0]
                tgt = bc['bval'].reshape(-1).to(device, torch.float32)
                if tgt.numel() != n:
                    raise ValueError(f'{tgt.numel()} target values for {n} boundary points')
            tb = te = 0
            if bop is not None:
                tb, te = add_terms(bop)
            bc_global.append(dict(cells=cidx.to(torch.int64), tgt=tgt, K=K, var=int(bc['var']), slot=slot, tb=tb, te=te,
                                  sign=sign, type_off=type_len[kind]))
            type_len[kind] += n
        self.type_len = [type_len[t] for t in self.bnd_types]
        max_len = max(self.type_len)
        self.n_eq = n_eq
        self.n_slots = n_eq + len(self.bnd_types)
        self.slot_len = [n0 * n1] * n_eq + [max_len] * len(self.bnd_types)        # global counts
        self.slot_lambda = _lambda_list(lambda_operator, n_eq, 'lambda_operator') + \
            _lambda_list(lambda_bound, len(self.bnd_types), 'lambda_bound')

        # ---- rows of this rank ---------------------------------------------------------------------------
        if len(self.fields) > 12:
            raise UnsupportedProblem('more than 12 distinct derivative fields')
        self.reach0 = max([order * (p - 1) for (_, axis, order) in self.fields if order > 0 and axis == 0] + [0])
        self.halo = 2 * self.reach0
        r0, r1 = slab_rows(n0, rank, world)
        if world > 1 and (r1 - r0) < max(self.halo, 1):
            raise UnsupportedProblem(f'{n0} grid rows are too few for {world} ranks (halo {self.halo})')
        e0, e1 = (max(0, r0 - self.halo), min(n0, r1 + self.halo)) if world > 1 else (0, n0)
        self.rows, self.ext = (r0, r1), (e0, e1)
        n_ext = e1 - e0
        self.shape = (n_var, r1 - r0, n1)                                        # the model tensor of this rank
        self.shape_ext = (n_var, n_ext, n1)
        self.n_cells = n0 * n1
        self.n_cells_local = (r1 - r0) * n1

        # coefficient buffers: the extended slab of every per-cell tensor
        for t in terms:
            if t[1] == 1:
                t[2] = t[2] * n_ext * n1                                         # buffer number -> float offset
        self.coeffs = (torch.cat([c[e0:e1].reshape(-1) for c in coefs_full]).contiguous() if coefs_full
                       else torch.zeros(1, device=device))

        # boundary rows owned by this rank, cells as flat indices into the extended slab
        bc_rows, cells, targets = [], [], []
        self.bc_layout = []                        # (type index, offset in type column, n) per condition (unsharded)
        cell_off = tgt_off = 0
        for b in bc_global:
            c = b['cells']
            i0 = c // n1
            own = (i0[:, 0] >= r0) & (i0[:, 0] < r1)
            if world > 1 and b['K'] > 1:
                inside = ((i0 >= e0) & (i0 < e1)).all(1)
                if not bool((inside | ~own).all()):
                    raise UnsupportedProblem('a periodic condition couples rows of different ranks')
            if world > 1 and b['te'] > b['tb']:
                # a boundary operator that differentiates along the sharded axis reaches into the neighbour's rows: the part
                # of its adjoint that falls there would be cut off (the neighbour never evaluates this row)
                reach = max([self.fields[f[1]][2] * (p - 1) for t in terms[b['tb']:b['te']] for f in factors[t[3]:t[4]]
                             if self.fields[f[1]][1] == 0 and self.fields[f[1]][2] > 0] + [0])
                if reach > 0:
                    rows_own = i0[own][:, 0]
                    near = ((r0 > 0) & (rows_own < r0 + reach)) | ((r1 < n0) & (rows_own >= r1 - reach))
                    if bool(near.any()):
                        raise UnsupportedProblem('a boundary operator with a derivative along grid axis 0 sits within its '
                                                 'stencil reach of a slab interface: its adjoint would cross ranks')
            c = c[own] - e0 * n1
            tgt = b['tgt'][own.to(b['tgt'].device)]
            n = int(c.shape[0])
            bc_rows.append((n, cell_off, tgt_off, b['K'], b['var'], b['slot'], b['tb'], b['te'], b['sign']))
            self.bc_layout.append((b['slot'], b['type_off'], n))
            cells.append(c.reshape(-1))
            targets.append(tgt)
            cell_off += n * b['K']
            tgt_off += n

        # ---- bands for every derivative field (on the extended slab along axis 0) -----------------------
        dims = (n_ext, n1)
        fld = np.zeros(len(self.fields), FIELD_DTYPE)
        band = [np.zeros(1, np.float32)]
        off = 1
        for q, (var, axis, order) in enumerate(self.fields):
            fld[q]['var'], fld[q]['axis'], fld[q]['order'] = var, axis, order
            if order > 0:
                bnd_arr, b, E = derivative_band(dims[axis], p, order, hs[axis])
                fld[q]['half_width'], fld[q]['n_edge'], fld[q]['coef_off'] = b, E, off
                band.append(bnd_arr)
                off += bnd_arr.size
        self.fld = fld
        self.band = np.concatenate(band).astype(np.float32)
        self.terms = np.array([tuple(t) for t in terms], dtype=TERM_DTYPE) if terms else np.zeros(0, TERM_DTYPE)
        self.factors = np.array(factors, dtype=FACTOR_DTYPE) if factors else np.zeros(0, FACTOR_DTYPE)
        self.bcs = np.zeros(len(bc_rows), _native.MAT_BC_DTYPE)
        for i, (n, co, to, K, var, slot, tb, te, sign) in enumerate(bc_rows):
            r = self.bcs[i]
            r['n_rows'], r['cell_off'], r['tgt_off'], r['K'], r['var'], r['slot'] = n, co, to, K, var, slot
            r['term_begin'], r['term_end'] = tb, te
            r['sign'][:] = sign
        self.cells = (torch.cat(cells) if cells else torch.zeros(0, dtype=torch.int64)).to(device, torch.int32).contiguous()
        self.targets = (torch.cat(targets) if targets else torch.zeros(0, device=device)).contiguous()
        self.n_bc_rows = int(self.targets.numel())


def open_peer(lib, handle, created: bool, world: int, group=None) -> bool:
    """Collective: exchanges the CUDA-IPC handles of the ranks' exchange blocks over torch.distributed and maps them
    (csrc/peer.cu).  True on every rank if every rank sits on the same host and could map every block; otherwise the
    object is destroyed everywhere and False is returned (the callers fall back to NCCL)."""
    import os
    import torch.distributed as dist
    mine = (C.c_char * 64)()
    ok = created and lib.tdb200_peer_handle(handle, mine) == 0
    gathered = [None] * world
    dist.all_gather_object(gathered, (bytes(mine.raw) if ok else None, os.uname().nodename), group=group)
    same_box = all(g[0] is not None and g[1] == gathered[0][1] for g in gathered)
    if same_box:
        same_box = lib.tdb200_peer_open(handle, b''.join(g[0] for g in gathered)) == 0
    flags = [None] * world
    dist.all_gather_object(flags, bool(same_box), group=group)               # all ranks take the same path
    if all(flags):
        return True
    if created:
        lib.tdb200_peer_destroy(handle)
    return False


def _robin_operator(bop: dict, var: int) -> dict:
    """`Bounds._apply_robin` (tedeous/eval.py:357-388) as a boundary operator: alpha * u + sum_beta beta * (the whole bop
    applied), alpha and the betas being the coefficients of bop's terms - the alpha * u term is counted again inside
    every beta term, reproduced.  Numeric alpha / beta (in mat mode the reference would call a callable
    one on cell positions, not on coordinates)."""
    labels = list(bop.keys())
    coeffs = [bop[k]['coeff'] for k in labels]
    for c in coeffs:
        if not isinstance(c, (int, float)) and not (isinstance(c, torch.Tensor) and c.numel() == 1):
            raise UnsupportedProblem('robin conditions in mat mode take numeric coefficients')
    alpha, betas = float(coeffs[0]), [float(c) for c in coeffs[1:]]
    dif = list(bop[labels[0]].keys())[1]
    out = {'robin:alpha*u': {'coeff': alpha, dif: [[None]], 'pow': [1], 'var': [var]}}
    for i, beta in enumerate(betas):
        for k in labels:
            t = dict(bop[k])
            t['coeff'] = beta * float(t['coeff'])
            out[f'robin:beta{i}*{k}'] = t
    return out


def exchange_halos(u: torch.Tensor, ir: MatIR, group=None, ext: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[n_var, n_local, n1] -> the extended slab [n_var, n_ext, n1]: the owned rows plus `halo` rows received from each
    neighbour (one all-gather of the edge rows; NCCL on GPUs, gloo in the CPU tests).  Single rank: returns `u` itself.
    If `ext` is given and `u` already is the view of its owned rows (MatPlan adopts the model tensor that way), only
    the halo rows move."""
    rank, world = ir.shard
    if world == 1:
        return u
    import torch.distributed as dist
    (r0, r1), (e0, e1) = ir.rows, ir.ext
    up, down, n = r0 - e0, e1 - r1, r1 - r0                     # halo rows above / below the owned block
    in_place = ext is not None and u.data_ptr() == ext[:, up:up + n].data_ptr() and u.shape[0] == 1
    if not in_place:
        ext = torch.empty(ir.shape_ext, dtype=u.dtype, device=u.device)
        ext[:, up:up + n] = u
    # one all-gather of every rank's first / last `halo` owned rows (world * 2 * halo rows: ~1 MB at 8 ranks of a
    # 4096-wide grid) instead of four point-to-point operations: a single collective launch per step
    h = ir.halo
    if h == 0:
        return ext
    mine = torch.stack([ext[:, up:up + h], ext[:, up + n - h:up + n]])          # [2, n_var, h, n1]
    allh = torch.empty((world,) + tuple(mine.shape), dtype=u.dtype, device=u.device)
    if u.is_cuda:
        dist.all_gather_into_tensor(allh, mine.contiguous(), group=group)
    else:
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine.contiguous(), group=group)
        allh = torch.stack(parts)
    if rank > 0:
        ext[:, 0:up] = allh[rank - 1, 1]                       # last rows of the neighbour above
    if rank < world - 1:
        ext[:, up + n:up + n + down] = allh[rank + 1, 0]       # first rows of the neighbour below
    return ext


class MatPlan:
    """Owns one tdb200_mat_plan (built from the `MatIR` of this rank)."""

    def __init__(self, grid: torch.Tensor, prepared_operator: List[dict], bconds: List[dict], model: torch.Tensor,
                 lambda_operator, lambda_bound, derivative_points: int = 2, shard=None, process_group=None,
                 per_cell_coeffs: bool = False):
        # 1-D grids ([1, N0], model [n_eq, N0]: the ODE examples, example_ODE_Legendre_matrix.py, example_LV_mat.py) run as
        # [N0, 1] grids with a dummy second axis - same memory, no derivative field along it
        self._lift = grid.dim() == 2 and model.dim() == 2
        if self._lift:
            if shard is not None and shard[1] > 1:
                raise UnsupportedProblem('1-D mat-mode grids are not sharded over ranks')
            model = model.detach().unsqueeze(-1)          # (MatIR lifts the grid the same way)
        if model.dim() != 3:
            raise UnsupportedProblem('the fused mat path supports 1-D and 2-D grids ([d, N0(, N1)], model [n_eq, N0(, N1)])')
        if model.dtype != torch.float32:
            raise UnsupportedProblem('mat-mode model must be float32')
        shard = (0, 1) if shard is None else shard
        self.lib = _native.load()
        self.device = model.device
        self.grid = grid
        self._pg = process_group
        ir = MatIR(grid, prepared_operator, bconds, int(model.shape[0]), model.device, lambda_operator, lambda_bound,
                   derivative_points, shard, per_cell_coeffs)
        self.ir = ir
        if tuple(model.shape) != ir.shape:
            raise UnsupportedProblem(f'rank {shard[0]} of {shard[1]} owns grid rows {ir.rows}: its model tensor must be '
                                     f'{ir.shape}, got {tuple(model.shape)}')
        self.shape, self.n_eq, self.n_slots = ir.shape, ir.n_eq, ir.n_slots
        self.bnd_types, self.type_len = ir.bnd_types, ir.type_len
        self.slot_len, self.slot_lambda = ir.slot_len, ir.slot_lambda
        self._bc_layout = ir.bc_layout
        n_var, n_ext, n1 = ir.shape_ext
        desc = _native.MatDesc(ir.n_eq, n_var, n_ext, n1, len(ir.fields))
        eb = np.array([r[0] for r in ir.eq_ranges], np.int32)
        ee = np.array([r[1] for r in ir.eq_ranges], np.int32)
        handle = C.c_void_p()
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        _native.check(self.lib.tdb200_mat_plan_create(
            C.byref(desc), _native.np_ptr(ir.fld), ir.band.size, _native.np_ptr(ir.band), _native.np_ptr(eb),
            _native.np_ptr(ee), len(ir.terms), _native.np_ptr(ir.terms), len(ir.factors), _native.np_ptr(ir.factors),
            dev_index, C.byref(handle)), 'tdb200_mat_plan_create')
        self.handle = handle
        self._coeffs = ir.coeffs
        _native.check(self.lib.tdb200_mat_plan_set_coeffs(handle, self._coeffs.data_ptr(), self._coeffs.numel()),
                      'tdb200_mat_plan_set_coeffs')
        self._bcs, self._cells, self._targets = ir.bcs, ir.cells, ir.targets
        if self._cells.numel() == 0:                      # keep valid device pointers for ranks without boundary rows
            self._cells = torch.zeros(1, dtype=torch.int32, device=self.device)
            self._targets = torch.zeros(1, dtype=torch.float32, device=self.device)
        self.n_bc_rows = ir.n_bc_rows
        self._push_bcs()
        if shard[1] > 1:
            _native.check(self.lib.tdb200_mat_plan_set_row_window(handle, ir.rows[0] - ir.ext[0], ir.rows[1] - ir.ext[0]),
                          'tdb200_mat_plan_set_row_window')
        self.out_size = int(self.lib.tdb200_mat_plan_out_size(handle))
        self.launches_per_call = int(self.lib.tdb200_mat_plan_launches_per_call(handle))
        self.kernel_kind = ('generic', 'register-tap', 'cross-vec4', 'cross-tma', 'cross-march')[int(self.lib.tdb200_mat_plan_kernel_kind(handle))]
        self.n_cells = ir.n_cells                          # global
        self.n_cells_local = ir.n_cells_local
        # Several ranks, one field: the model tensor is adopted as the owned-row view of a persistent extended slab, so
        # a step moves only the halo rows (optimisers update the view in place; `model.data` is re-pointed once here)
        self._ext = None
        if shard[1] > 1 and n_var == 1:
            up = ir.rows[0] - ir.ext[0]
            self._ext = torch.zeros(ir.shape_ext, dtype=torch.float32, device=self.device)
            self._ext[:, up:up + ir.shape[1]] = model.detach()
            model.data = self._ext[:, up:up + ir.shape[1]]
        self._peer = None
        self._peer_loss_inline = False
        self._setup_peer()

    def _setup_peer(self):
        """Peer-memory exchange (csrc/peer.cu) for the halo rows and the loss terms when the ranks are the GPUs of one
        box: CUDA-IPC handles of the per-rank exchange blocks travel once over torch.distributed; every step then moves
        its 2 x halo rows and its [2 + n_slots] loss terms with two small kernels of the library instead of two NCCL
        collectives.  TDB200_MAT_COLLECTIVE=nccl keeps the library collectives (also the fallback when the handles
        cannot be opened, e.g. ranks on different hosts)."""
        import os
        ir = self.ir
        rank, world = ir.shard
        if (world == 1 or self._ext is None or ir.halo == 0 or not self.device.type == 'cuda'
                or os.environ.get('TDB200_MAT_COLLECTIVE', 'peer') != 'peer' or self.out_size > 64):
            return
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            return
        n_var, n_ext, n1 = ir.shape_ext
        if (ir.halo * n1) % 4:
            return
        handle = C.c_void_p()
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        ok = self.lib.tdb200_peer_create(rank, world, n_var * ir.halo * n1, 0, dev_index, C.byref(handle)) == 0
        if open_peer(self.lib, handle, ok, world, self._pg):
            self._peer = handle
            # the finalizing block of the boundary kernel exchanges the loss terms itself where the schedule allows it
            flag = C.c_int32(0)
            _native.check(self.lib.tdb200_mat_plan_set_peer(self.handle, handle, C.byref(flag)), 'tdb200_mat_plan_set_peer')
            self._peer_loss_inline = bool(flag.value)

    def set_cell_weights(self, w: Optional[torch.Tensor]):
        """Causal loss in mat mode (tedeous/losses.py:137-182 with `n_t = grid.shape[1]`, solution.py:58-60): the operator part
        of the loss becomes sum_cells w * sum_eq res^2 / N with no-grad weights w [N0, N1].  Every equation term reads its
        coefficient from a per-cell buffer (`per_cell_coeffs`), and res is linear in those coefficients, so scaling them
        by sqrt(w) makes the unchanged kernels evaluate exactly that loss and its gradient.  None restores the plain
        coefficients."""
        nb = self.ir.n_eq_buffers
        if nb == 0:
            raise UnsupportedProblem('set_cell_weights needs a plan built with per_cell_coeffs=True')
        n_cells = self.ir.shape_ext[1] * self.ir.shape_ext[2]
        if getattr(self, '_coeffs_base', None) is None:
            self._coeffs_base = self._coeffs.clone()
        if w is None:
            self._coeffs.copy_(self._coeffs_base)
            return
        sw = torch.sqrt(w.detach().to(self.device, torch.float32)).reshape(-1)
        if sw.numel() != n_cells:
            raise ValueError('one weight per grid cell expected')
        view, base = self._coeffs[:nb * n_cells].view(nb, n_cells), self._coeffs_base[:nb * n_cells].view(nb, n_cells)
        torch.mul(base, sw, out=view)

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
        if self._lift and tuple(u.shape) == self.shape[:2]:
            return
        if tuple(u.shape) != self.shape or u.dtype != torch.float32 or not u.is_cuda or not u.is_contiguous():
            raise RuntimeError(f'mat-mode model must be a contiguous float32 CUDA tensor of shape {self.shape}')

    def loss_grad_ext(self, ue: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """Kernel launch on this rank's extended slab (no communication): -> (partial out [2 + n_slots] of this rank,
        d loss / d u of the owned rows)."""
        if tuple(ue.shape) != self.ir.shape_ext or ue.dtype != torch.float32 or not ue.is_cuda or not ue.is_contiguous():
            raise RuntimeError(f'extended slab must be a contiguous float32 CUDA tensor of shape {self.ir.shape_ext}')
        out = torch.empty(self.out_size, dtype=torch.float32, device=self.device)
        grad = torch.empty(self.ir.shape_ext, dtype=torch.float32, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _native.check(self.lib.tdb200_mat_loss_grad(self.handle, ue.data_ptr(), grad.data_ptr(), out.data_ptr(),
                                                    stream), 'tdb200_mat_loss_grad')
        if self.ir.shard[1] > 1:
            up = self.ir.rows[0] - self.ir.ext[0]
            grad = grad[:, up:up + self.shape[1]].contiguous()
        return out, grad

    def set_timing(self, on: bool):
        """Measurement aid: CUDA events around the stencil-kernel launch of every following eager call."""
        _native.check(self.lib.tdb200_mat_plan_set_timing(self.handle, 1 if on else 0), 'tdb200_mat_plan_set_timing')

    def time_stencil(self, ue: torch.Tensor, iters: int = 10) -> float:
        """Mean duration (ms) of `iters` back-to-back launches of the stencil kernel alone on the extended slab."""
        import ctypes
        grad = torch.empty(self.ir.shape_ext, dtype=torch.float32, device=self.device)
        ms = ctypes.c_float(0.0)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _native.check(self.lib.tdb200_mat_time_stencil(self.handle, ue.data_ptr(), grad.data_ptr(), int(iters),
                                                       ctypes.byref(ms), stream), 'tdb200_mat_time_stencil')
        return float(ms.value)

    def stencil_ms(self) -> float:
        """Duration of the stencil kernel of the last eager call made with timing on (waits for it)."""
        import ctypes
        ms = ctypes.c_float(0.0)
        _native.check(self.lib.tdb200_mat_plan_stencil_ms(self.handle, ctypes.byref(ms)), 'tdb200_mat_plan_stencil_ms')
        return float(ms.value)

    def loss_grad_raw(self, u: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """-> (out [2 + n_slots] summed over ranks, d loss / d u of this rank's rows).  Several ranks: halo rows from
        the neighbours (point-to-point), then one all-reduce of the loss terms; the gradient stays sharded."""
        self._check_model(u)
        if self._lift and u.dim() == 2:
            out, grad = self.loss_grad_ext(u.unsqueeze(-1))
            return out, grad.squeeze(-1)
        if self._peer is not None and u.data_ptr() == self._own_view_ptr():
            # halo rows and loss terms over peer memory (two kernels of the library, csrc/peer.cu)
            ir = self.ir
            n_var, n_ext, n1 = ir.shape_ext
            up, n, h = ir.rows[0] - ir.ext[0], ir.rows[1] - ir.rows[0], ir.halo
            stream = torch.cuda.current_stream(self.device).cuda_stream
            _native.check(self.lib.tdb200_peer_halo(self._peer, self._ext.data_ptr(), n_ext * n1, n_var, h * n1, up * n1,
                                                    (up + n - h) * n1, 0, (up + n) * n1, stream), 'tdb200_peer_halo')
            out, grad = self.loss_grad_ext(self._ext)
            if not self._peer_loss_inline:
                _native.check(self.lib.tdb200_peer_allreduce(self._peer, out.data_ptr(), self.out_size, stream),
                              'tdb200_peer_allreduce')
            return out, grad
        out, grad = self.loss_grad_ext(exchange_halos(u, self.ir, self._pg, self._ext))
        if self.ir.shard[1] > 1:
            import torch.distributed as dist
            dist.all_reduce(out, op=dist.ReduceOp.SUM, group=self._pg)
        return out, grad

    def _own_view_ptr(self) -> int:
        up = self.ir.rows[0] - self.ir.ext[0]
        return self._ext[:, up:up + self.shape[1]].data_ptr()

    def peer_error(self) -> bool:
        """True if a peer-memory wait timed out (a rank fell out of step); synchronises."""
        if self._peer is None:
            return False
        err = C.c_int32(0)
        _native.check(self.lib.tdb200_peer_error(self._peer, C.byref(err)), 'tdb200_peer_error')
        return bool(err.value)

    def capture(self, u: torch.Tensor):
        """CUDA graph of one step - halo exchange, the two kernel launches, all-reduce of the loss terms - for training
        loops that are launch bound (a step is ~90 us of GPU work).  -> (replay callable, out, grad): `out` / `grad`
        are static tensors refreshed by every replay; `u` must keep its storage (in-place optimiser updates do)."""
        self._check_model(u)
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(3):                          # warm-up outside the capture (attributes, tensor maps, NCCL)
                self.loss_grad_raw(u)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out, grad = self.loss_grad_raw(u)
        return graph.replay, out, grad

    def eval_fields(self, u: torch.Tensor):
        self._check_model(u)
        if self.ir.shard[1] > 1:
            raise UnsupportedProblem('per-point fields are not gathered across ranks')
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
        if self.shape[0] > 1:
            # the reference forms every term as ones_like(model) * field (derivative.py:306): with n_var fields in the
            # model the residual of an equation comes out n_var times, [n_var * N, n_eq] - reproduced for `Solution.op`
            op = op.repeat(self.shape[0], 1)
        return op, bval, tval

    def __del__(self):
        try:
            if getattr(self, '_peer', None):
                self.lib.tdb200_peer_destroy(self._peer)
                self._peer = None
            if getattr(self, 'handle', None):
                self.lib.tdb200_mat_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass
