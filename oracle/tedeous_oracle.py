"""ORACLE - test infrastructure, not product code.

A self-contained CPU restatement (torch CPU tensors + torch.autograd, fp32 or fp64) of the reference's
residual-loss hot path, TEDEouS v0.4.11:

    Solution.evaluate           tedeous/solution.py:129-168
    Operator / Bounds           tedeous/eval.py:55-87, 143-193, 235-461
    Derivative_NN/_autograd/_mat tedeous/derivative.py:18-323
    Finite_diffs                tedeous/finite_diffs.py:9-268
    Points_type                 tedeous/points_type.py:40-157
    Losses                      tedeous/losses.py:38-263 (default, causal and weak-form losses)
    integration / weak residual tedeous/eval.py:13-52, 195-221
    lambda_prepare / unify      tedeous/input_preprocessing.py:14-81, 239-264, 319-408, 553-575

The arithmetic of the reference lives in torch (requirements.txt:6 `torch >= 2.0`, unpinned; torch
2.11.0+cu128 here): `model(grid)`, `torch.autograd.grad`, `torch.roll`, `scipy.linalg.solve`.  The oracle
keeps exactly those call patterns (one MLP forward per finite-difference shift in NN mode, nested
`autograd.grad(create_graph=True)` in autograd mode, rolled one-sided differences in mat mode) so it is
also the "port" CPU baseline `bench.py` times.

Pinning: the reference ships no tests for this path, so the oracle is pinned against outputs of
the reference itself, generated in the build container by `tests/golden/make_golden.py` (imports
/root/reference) and committed as tests/golden/*.npz; `tests/test_oracle_golden.py` checks every fixture.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module.  Nothing under torch_de_solver_b200/ does.

Deliberate deviation (documented in DESIGN.md): NN-mode boundary-operator values are returned in `bnd`
order; the reference concatenates them per point-type subset in Python-set order, which depends on
PYTHONHASHSEED.  The golden generator asserts the two orders agree for its fixtures.
"""
from __future__ import annotations

from copy import deepcopy
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np
import torch


# ------------------------------------------------------------------------------------------------
# input_preprocessing.py:14-81
# ------------------------------------------------------------------------------------------------
def lambda_prepare(val, lam):
    if isinstance(lam, torch.Tensor):
        return lam
    if isinstance(lam, (int, float)):
        lam = torch.ones(val.shape[-1]) * lam
    else:
        lam = torch.tensor(lam)
    return lam.reshape(1, -1)


def unify(eq: dict) -> dict:
    for label in eq:
        t = eq[label]
        dif = list(t.keys())[1]
        scalar = isinstance(t['pow'], (int, float)) or callable(t['pow'])
        if 'var' not in t:
            if scalar:
                t[dif], t['pow'], t['var'] = [t[dif]], [t['pow']], [0]
            else:
                t['var'] = [0] * len(t['pow'])
        elif scalar:
            t[dif], t['pow'], t['var'] = [t[dif]], [t['pow']], [t['var']]
    return eq


# ------------------------------------------------------------------------------------------------
# finite_diffs.py: recursive shift / sign construction
# ------------------------------------------------------------------------------------------------
def fd_scheme(term: list, nvars: int, ptype: str, label: str, h: float):
    """-> (shifts, signs) exactly as Finite_diffs(term, nvars, ptype).scheme_choose(label, h)."""
    if term == [None]:
        return [None], [1]
    dirs = ['central'] * len(term) if ptype == 'central' else [ptype[a] for a in term]
    shifts, signs = [[0] * nvars], [1]
    for ax, dr in zip(term, dirs):
        ns, ng = [], []
        for s, g in zip(shifts, signs):
            def sh(delta):
                c = list(s)
                c[ax] += delta
                return c
            if label == '1' or dr == 'central':
                # First_order_scheme._finite_diff_shift (finite_diffs.py:37-57) + sign_order (86-117)
                plus, minus = {'central': (1, -1), 'f': (1, 0), 'b': (0, -1)}[dr]
                ns += [sh(plus), sh(minus)]
                w = 1 / (2 * h) if dr == 'central' else 1 / h
                ng += [g * w, -g * w] if dr == 'central' else [g / h, -g / h]
            else:
                # Second_order_scheme (finite_diffs.py:146-223): points x+-2h, x+-h, x with 3, -4, 1
                sgn = 1 if dr == 'f' else -1
                ns += [sh(2 * sgn), sh(sgn), sh(0)]
                ng += [sgn * 3 * (1 / (2 * h)) * g, -sgn * 4 * (1 / (2 * h)) * g, sgn * (1 / (2 * h)) * g]
        shifts, signs = ns, ng
    return shifts, signs


# ------------------------------------------------------------------------------------------------
# points_type.py:40-157 (Delaunay hull test, vectorised over points)
# ------------------------------------------------------------------------------------------------
def point_types(grid: torch.Tensor) -> List[str]:
    n, d = grid.shape
    if d == 1:
        return ['central'] * n
    from scipy.spatial import Delaunay
    hull = Delaunay(grid.cpu().numpy())
    flags = []
    for axis in range(d):
        for direction in range(2):
            sh = grid.clone()
            sh[:, axis] = grid[:, axis] + (-1) ** direction * 0.0001
            flags.append(hull.find_simplex(sh.cpu().numpy()) >= 0)
    flags = np.array(flags).T                               # [n, 2d]: (+a, -a) pairs
    out = []
    for row in flags:
        if row.all():
            out.append('central')
        else:
            out.append(''.join('f' if row[2 * a] else 'b' for a in range(d)))
    return out


# ------------------------------------------------------------------------------------------------
# derivative.py
# ------------------------------------------------------------------------------------------------
def _coeff(term, pts):
    c = term['coeff']
    if callable(c) and not isinstance(c, torch.Tensor):
        return c(pts).reshape(-1, 1)
    if isinstance(c, torch.Tensor) and c.dim() > 0 and c.numel() > 1:
        return c.reshape(-1, 1)
    return c


def _apply_pow(acc, val, pw):
    if callable(pw):
        return pw(acc * val)
    return acc * val ** pw


def nn_autograd(model, points, var, axis):
    """derivative.py:73-98."""
    points.requires_grad = True
    fi = model(points)[:, var].sum(0)
    for ax in axis:
        grads, = torch.autograd.grad(fi, points, create_graph=True)
        fi = grads[:, ax].sum()
    return grads[:, axis[-1]].reshape(-1, 1)


def term_autograd(model, term, pts):
    """Derivative_autograd.take_derivative (derivative.py:100-132)."""
    dif = list(term.keys())[1]
    acc = 1.
    for j, der in enumerate(term[dif]):
        if der == [None]:
            val = model(pts)[:, term['var'][j]].reshape(-1, 1)
        else:
            val = nn_autograd(model, pts, term['var'][j], der)
        acc = _apply_pow(acc, val, term['pow'][j])
    return _coeff(term, pts) * acc


def term_nn(model, term, pts, ptype, h, inner_order, boundary_order, coeff_override=None):
    """Derivative_NN.take_derivative on shifted copies of `pts` (derivative.py:30-58 with the grids of
    input_preprocessing.py:208-264: one `model(grid + shift*h)` forward per stencil entry)."""
    dif = list(term.keys())[1]
    label = inner_order if ptype == 'central' else boundary_order
    acc = 1.
    for j, der in enumerate(term[dif]):
        shifts, signs = fd_scheme(der, pts.shape[1], ptype, label, h)
        total = 0.
        for s, g in zip(shifts, signs):
            p = pts
            if s is not None:
                for a, mult in enumerate(s):
                    q = p.clone()                    # points_type.py:22-37 (applied even for mult == 0)
                    q[:, a] = p[:, a] + mult * h
                    p = q
            total = total + model(p)[:, term['var'][j]].reshape(-1, 1) * g
        acc = _apply_pow(acc, total, term['pow'][j])
    c = coeff_override if coeff_override is not None else _coeff(term, pts)
    return c * acc


class MatDerivative:
    """Derivative_mat (derivative.py:135-323)."""

    def __init__(self, derivative_points: int):
        from scipy import linalg
        p = derivative_points
        self.backward = list(range(-p + 1, 1))
        self.farward = list(range(p))

        def alpha(labels):
            lab = np.array(labels)
            A = np.array([lab ** i for i in range(len(labels))])
            b = np.zeros_like(lab)
            b[1] = 1
            return linalg.solve(A, b)
        self.alpha_backward, self.alpha_farward = alpha(self.backward), alpha(self.farward)
        self.back = [int(0 - i) for i in range(1, p)]
        self.farw = [int(i) for i in range(p - 1)]

    @staticmethod
    def step_h(grid):
        flat = torch.vstack([grid[i].reshape(-1) for i in range(grid.shape[0])]).T.float()
        out = []
        for i in range(flat.shape[-1]):
            u = torch.unique(flat[:, i])
            out.append(abs(u[1] - u[0]))
        return out

    def d1(self, u, h, axis):
        if u.dim() == 1 or u.shape[0] == 1:
            shape = u.shape
            u = u.reshape(-1)
            db = sum(torch.roll(u, -sb) * a for sb, a in zip(self.backward, self.alpha_backward))
            df = sum(torch.roll(u, -sf) * a for sf, a in zip(self.farward, self.alpha_farward))
            du = (db + df) / (2 * h)
            du[self.back] = db[self.back] / h
            du[self.farw] = df[self.farw] / h
            return du.reshape(shape)
        pos = u.dim() - 1
        u = torch.transpose(u, pos, axis)
        db = sum(torch.roll(u, -sb) * a for sb, a in zip(self.backward, self.alpha_backward))
        df = sum(torch.roll(u, -sf) * a for sf, a in zip(self.farward, self.alpha_farward))
        du = (db + df) / (2 * h)
        du[..., self.back] = db[..., self.back] / h
        du[..., self.farw] = df[..., self.farw] / h
        return torch.transpose(du, pos, axis)

    def term(self, model, term, grid):
        dif = list(term.keys())[1]
        acc = torch.zeros_like(model) + 1
        for j, scheme in enumerate(term[dif]):
            prod = model[term['var'][j]]
            if scheme != [None]:
                for axis in scheme:
                    if axis is None:
                        continue
                    prod = self.d1(prod, self.step_h(grid)[axis], axis)
            acc = _apply_pow(acc, prod, term['pow'][j])
        c = term['coeff']
        return (c(grid) if callable(c) and not isinstance(c, torch.Tensor) else c) * acc


# ------------------------------------------------------------------------------------------------
def integration(func, grid, power=2):
    """eval.py:13-52, restated loop for loop: trapezoid rule along the last grid column inside runs of equal values
    of the column before it (one run when the grid has a single column), integrand raised to `power`."""
    column = -1 if grid.shape[-1] == 1 else -2
    marker = grid[0][column]
    index, result, u = [0], [], 0.
    for i in range(1, len(grid)):
        if grid[i][column] == marker or column == -1:
            u = u + (grid[i][-1] - grid[i - 1][-1]).item() * (func[i] ** power + func[i - 1] ** power) / 2
        else:
            result.append(u)
            marker = grid[i][column]
            index.append(i)
            u = 0.
    if column == -1:
        return u, 0.
    result.append(u)
    return result, grid[index, :-1]


# the problem object: Solution.evaluate
# ------------------------------------------------------------------------------------------------
class OracleSolution:
    """evaluate() -> (loss [1], loss_normalized [1]); also sets op, bval, true_bval, bval_keys,
    bval_length like tedeous.solution.Solution."""

    def __init__(self, grid, equations, bconds, model, mode, lambda_operator, lambda_bound, h=0.001,
                 inner_order='1', boundary_order='2', derivative_points=2, tol=0, weak_form=None):
        self.grid, self.model, self.mode = grid, model, mode
        self.weak_form = weak_form
        self.h, self.inner_order, self.boundary_order = h, inner_order, boundary_order
        self.lambda_operator, self.lambda_bound, self.tol = lambda_operator, lambda_bound, tol
        eqs = equations if isinstance(equations, list) else [equations]
        self.equations = [unify(deepcopy_keep_params(e)) for e in eqs]
        self.bconds = []
        for bc in bconds:
            bc = dict(bc)
            if bc['bop'] is not None:
                bc['bop'] = unify(deepcopy_keep_params(bc['bop']))
            self.bconds.append(bc)
        if mode == 'NN':
            self.types = point_types(grid)
            self.central = torch.tensor([t == 'central' for t in self.types])
            self.grid_central = grid[self.central]
            self.n_t = len(self.grid_central[:, 0].unique())
        elif mode == 'autograd':
            self.n_t = len(grid[:, 0].unique())
        else:
            self.n_t = grid.shape[1]
            self.mat = MatDerivative(derivative_points)
            self._mat_positions()

    # -- operator (eval.py:143-193) ------------------------------------------------------------------
    def _apply_operator(self, eq, pts, ptype='central'):
        total = None
        for label in eq:
            term = eq[label]
            if self.mode == 'NN':
                coeff = None
                c = term['coeff']
                if isinstance(c, torch.Tensor) and c.numel() == self.grid.shape[0] and c.numel() > 1 \
                        and pts is self.grid_central:
                    coeff = c.reshape(-1)[self.central].reshape(-1, 1)     # input_preprocessing.py:287-290
                v = term_nn(self.model, term, pts, ptype, self.h, self.inner_order, self.boundary_order, coeff)
            elif self.mode == 'autograd':
                v = term_autograd(self.model, term, pts)
            else:
                v = self.mat.term(self.model, term, self.grid)
            total = v if total is None else total + v
        return total

    def operator_compute(self):
        pts = self.grid_central if self.mode == 'NN' else self.grid
        cols = [self._apply_operator(eq, pts).reshape(-1, 1) for eq in self.equations]
        return cols[0] if len(cols) == 1 else torch.cat(cols, 1)

    # -- boundary (eval.py:283-461) ------------------------------------------------------------------
    def _mat_positions(self):
        """Equation_mat._point_position (input_preprocessing.py:553-575): isclose per axis."""
        def pos(bnd):
            out = []
            for pt in bnd:
                mask = torch.ones_like(self.grid[0], dtype=torch.bool)
                for a in range(self.grid.shape[0]):
                    mask &= torch.isclose(pt[a].float(), self.grid[a].float())
                out.append(torch.where(mask))
            return out
        for bc in self.bconds:
            bc['pos'] = [pos(b) for b in bc['bnd']] if bc['type'] == 'periodic' else pos(bc['bnd'])

    def _dirichlet(self, bnd, var, pos=None):
        if self.mode == 'mat':
            return torch.cat([self.model[var][p] for p in pos]).reshape(-1, 1)
        return self.model(bnd)[:, var].reshape(-1, 1)

    def _neumann(self, bnd, bop, pos=None):
        if self.mode == 'autograd':
            return self._apply_operator(bop, bnd)
        if self.mode == 'mat':
            var = bop[list(bop.keys())[0]]['var'][0]
            field = self._apply_operator(bop, None)
            return torch.cat([field[var][p] for p in pos]).reshape(-1, 1)
        # NN: one-sided stencils per point type (input_preprocessing.py:371-408), returned in bnd order
        from collections import OrderedDict
        types = self._bnd_types(bnd)
        out = torch.zeros(bnd.shape[0], 1, dtype=bnd.dtype)
        groups = OrderedDict()
        for i, t in enumerate(types):
            groups.setdefault(t, []).append(i)
        for t, idx in groups.items():
            out[idx] = self._apply_operator(bop, bnd[idx], t)
        return out

    def _bnd_types(self, bnd):
        types = []
        for b in bnd:                                         # points_type.py:141-150: exact match
            hit = torch.where((self.grid == b).all(dim=1))[0]
            types.append(self.types[int(hit[0])])
        return types

    def _bc_value(self, bc):
        kind, bnd, bop, var, pos = bc['type'], bc['bnd'], bc['bop'], bc['var'], bc.get('pos')
        if kind == 'dirichlet' or (kind == 'data' and bop is None):
            return self._dirichlet(bnd, var, pos)
        if kind in ('operator', 'data'):
            return self._neumann(bnd, bop, pos)
        if kind == 'periodic':
            f = (lambda b, p: self._dirichlet(b, var, p)) if bop is None else (lambda b, p: self._neumann(b, bop, p))
            val = f(bnd[0], pos[0] if pos else None).reshape(-1, 1).clone()
            for i in range(1, len(bnd)):
                val = val - f(bnd[i], pos[i] if pos else None).reshape(-1, 1)
            return val
        if kind == 'robin':                                   # eval.py:357-388
            coeffs = [bop[k]['coeff'] for k in bop]
            alpha, betas = coeffs[0], coeffs[1:]
            val = alpha * self._dirichlet(bnd, var, pos)
            for beta in betas:
                b = beta(bnd) if callable(beta) else beta
                val = val + b * self._neumann(bnd, bop, pos)
            return val
        raise ValueError(kind)

    def apply_bcs(self):
        vals: Dict[str, torch.Tensor] = {}
        true: Dict[str, torch.Tensor] = {}
        for bc in self.bconds:
            v, t = self._bc_value(bc).reshape(-1), bc['bval'].reshape(-1)
            k = bc['type']
            vals[k] = torch.cat((vals[k], v)) if k in vals else v
            true[k] = torch.cat((true[k], t)) if k in true else t
        keys = list(vals.keys())
        max_len = max(len(v) for v in vals.values())

        def pad(x):                                           # utils.py:257-290 via eval.py:55-87
            return torch.nn.functional.pad(x, (0, max_len - x.shape[-1]), value=0.)
        bval = torch.hstack([pad(vals[k]).reshape(-1, 1) for k in keys])
        tval = torch.hstack([pad(true[k]).reshape(-1, 1) for k in keys])
        return bval, tval, keys, [len(vals[k]) for k in keys]

    # -- loss (losses.py:84-182) -----------------------------------------------------------------------
    def _weak_operator(self, op):
        """eval.py:195-221: every equation column times the test functions, then one `integration` pass per grid
        column (each pass squares its integrand again - the reference's default power=2)."""
        pts = self.grid_central if self.mode == 'NN' else self.grid
        sols = []
        for i in range(op.shape[-1]):
            sol = op[:, i]
            for func in self.weak_form:
                sol = sol * func(pts).reshape(-1)
            g = torch.clone(pts)
            for _ in range(pts.shape[-1]):
                sol, g = integration(sol, g)
            sols.append(sol.reshape(-1, 1))
        return sols[0] if len(sols) == 1 else torch.cat(sols).reshape(1, -1)

    def evaluate(self):
        self.op = self.operator_compute()
        if self.weak_form not in (None, []):                      # losses.py:184-228
            self.op = self._weak_operator(self.op)
            self.bval, self.true_bval, self.bval_keys, self.bval_length = self.apply_bcs()
            dtype = self.op.dtype
            lam_op = lambda_prepare(self.op, self.lambda_operator).to(dtype)
            lam_b = lambda_prepare(self.bval, self.lambda_bound).to(dtype)
            bdiff = torch.mean((self.bval - self.true_bval) ** 2, 0)
            self.loss = self.op @ lam_op.T + bdiff @ lam_b.T
            with torch.no_grad():
                self.loss_normalized = self.op @ torch.ones_like(lam_op).T + bdiff @ torch.ones_like(lam_b).T
            return self.loss, self.loss_normalized
        self.bval, self.true_bval, self.bval_keys, self.bval_length = self.apply_bcs()
        dtype = self.op.dtype
        lam_op = lambda_prepare(self.op, self.lambda_operator).to(dtype)
        lam_b = lambda_prepare(self.bval, self.lambda_bound).to(dtype)
        bdiff = torch.mean((self.bval - self.true_bval) ** 2, 0)
        loss_bnd = bdiff @ lam_b.T
        if self.tol != 0:                                     # causal loss, losses.py:137-182
            res = torch.sum(self.op ** 2, dim=1).reshape(self.n_t, -1)
            m = torch.triu(torch.ones((self.n_t, self.n_t), dtype=res.dtype), diagonal=1).T
            with torch.no_grad():
                w = torch.exp(-self.tol * (m @ res))
            loss_op = torch.mean(w * res)
            self.loss = loss_op + loss_bnd
            with torch.no_grad():
                self.loss_normalized = loss_op + torch.ones(1, bdiff.shape[0], dtype=dtype) @ bdiff
            return self.loss, self.loss_normalized
        self.op_mse = torch.mean((self.op - torch.zeros(self.op.shape)) ** 2, 0)
        self.bval_mse = bdiff
        self.loss = self.op_mse @ lam_op.T + loss_bnd
        with torch.no_grad():
            self.loss_normalized = self.op_mse @ torch.ones_like(lam_op).T + bdiff @ torch.ones_like(lam_b).T
        return self.loss, self.loss_normalized


def deepcopy_keep_params(eq: dict) -> dict:
    """deepcopy of the term dicts that keeps nn.Parameter / tensor coefficients by reference
    (Solution._operator_coeff, solution.py:90-107)."""
    out = {}
    for label, term in eq.items():
        t = {}
        for k, v in term.items():
            t[k] = v if isinstance(v, torch.Tensor) or callable(v) else deepcopy(v)
        out[label] = t
    return out


def loss_and_grad(sol: OracleSolution, params: List[torch.Tensor]):
    """One optimiser-step evaluation: loss, loss_normalized and d loss / d params (closure.py:49-64)."""
    for p in params:
        p.grad = None
    loss, loss_n = sol.evaluate()
    loss.backward()
    return loss.detach(), loss_n.detach(), [p.grad.detach().clone() for p in params]
