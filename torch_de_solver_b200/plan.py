"""Lowering of a TEDEouS problem (equations + conditions + point sets) to the flat IR the CUDA kernels read.

The reference evaluates the loss by walking Python dicts every step (eval.py:143-193, 433-461,
derivative.py:30-132).  Here the walk happens once, at `Model.compile` time, and produces

* a **jet spec** per point set: which pure partial derivatives d^k/dx_a^k (k <= 4) the operators need.  The
  kernel propagates these as Taylor-mode channels through the tanh MLP (channel 0 = value);
* a **term table**: every operator is `sum_t coeff_t * prod_f chan[f]^pow_f` over (variable, channel) pairs;
  coeff is an immediate, a per-row buffer (callable / tensor coefficients, evaluated once) or a trainable
  scalar (`nn.Parameter`, inverse problems);
* a list of **segments**.  A segment is a set of residual rows that share operators and layout.  Each row
  is a *group* of K evaluation points; the group's channel values are a fixed linear combination
  V[m] = sum_{k,c} comb[m][k*J + c] * jet_c(x_k) of the per-point jets ("identity" for K = 1).  That one
  mechanism covers interior residuals (K = 1), Dirichlet/data rows (K = 1, value only), periodic rows
  (K = 2, V = u(x0) - u(x1)) and NN-mode one-sided finite-difference boundary operators (K = 3..5 shifted
  points, V = literal stencil - SURVEY B.1 q3);
* **loss slots**: slot s < n_eq is equation s (denominator = number of interior rows), the following
  slots are the boundary types in order of first appearance, all divided by the *longest* type
  (`dict_to_matrix` zero-padding, eval.py:55-87).

All arrays are numpy structured arrays whose dtypes mirror the C structs in include/tdb200.h.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .finite_diffs import Finite_diffs
from .points_type import Points_type

MAX_LAYERS = 16
MAX_DIRS = 4
MAX_ORDER = 4
MAX_COLS = 8
MAX_K = 32
MAX_M = 16
MAX_J = 16
ROWS_PER_TILE = 128          # (points x jet channels) rows a CTA tile holds; must match the kernel

SEGMENT_DTYPE = np.dtype([
    ('n_groups', '<i8'), ('pts_off', '<i8'), ('tgt_off', '<i8'), ('field_off', '<i8'),
    ('K', '<i4'), ('M', '<i4'), ('n_cols', '<i4'), ('n_dirs', '<i4'),
    ('dir_axis', '<i4', (MAX_DIRS,)), ('dir_order', '<i4', (MAX_DIRS,)),
    ('col_term_begin', '<i4', (MAX_COLS,)), ('col_term_end', '<i4', (MAX_COLS,)),
    ('col_slot', '<i4', (MAX_COLS,)),
    ('identity', '<i4'), ('comb_off', '<i4'),
    ('dir_vec', '<f4', (MAX_DIRS, 4)),
], align=True)

TERM_DTYPE = np.dtype([
    ('coeff', '<f4'), ('kind', '<i4'), ('idx', '<i8'), ('fac_begin', '<i4'), ('fac_end', '<i4'),
], align=True)

FACTOR_DTYPE = np.dtype([
    ('var', '<i4'), ('chan', '<i4'), ('pow', '<f4'), ('ipow', '<i4'),
], align=True)

COEFF_CONST, COEFF_BUFFER, COEFF_PARAM = 0, 1, 2


class UnsupportedProblem(NotImplementedError):
    """Raised for problem features the fused CUDA path does not implement (no silent CPU fallback)."""


# ----------------------------------------------------------------------------------------------------
# network description
# ----------------------------------------------------------------------------------------------------
@dataclass
class NetSpec:
    widths: List[int]                      # [d, w1, ..., n_out]
    linears: List[torch.nn.Linear]
    coeff_params: List[torch.nn.Parameter] = field(default_factory=list)

    @property
    def n_layers(self):
        return len(self.linears)

    @property
    def n_net_params(self):
        return sum(l.weight.numel() + l.bias.numel() for l in self.linears)

    def param_tensors(self) -> List[torch.Tensor]:
        out = []
        for l in self.linears:
            out += [l.weight, l.bias]
        return out + list(self.coeff_params)


def net_spec(model: torch.nn.Module) -> NetSpec:
    """Accepts `Sequential(Linear, Tanh, ..., Linear)` (every BASELINE config)."""
    if not isinstance(model, torch.nn.Sequential):
        raise UnsupportedProblem('the fused path supports torch.nn.Sequential(Linear, Tanh, ..., Linear) '
                                 f'only, got {type(model).__name__}')
    mods = list(model.children())
    if len(mods) < 3 or len(mods) % 2 == 0:
        raise UnsupportedProblem('expected Linear, Tanh, ..., Linear (odd number of modules, >= 3)')
    linears = []
    for i, m in enumerate(mods):
        if i % 2 == 0:
            if not isinstance(m, torch.nn.Linear) or m.bias is None:
                raise UnsupportedProblem(f'module {i} must be a Linear with bias, got {m}')
            linears.append(m)
        elif not isinstance(m, torch.nn.Tanh):
            raise UnsupportedProblem(f'module {i} must be Tanh, got {m}')
    if len(linears) > MAX_LAYERS:
        raise UnsupportedProblem(f'at most {MAX_LAYERS} Linear layers')
    widths = [linears[0].in_features] + [l.out_features for l in linears]
    for a, b in zip(linears[:-1], linears[1:]):
        if a.out_features != b.in_features:
            raise UnsupportedProblem('inconsistent layer widths')
    if max(widths[1:-1]) > 128:
        raise UnsupportedProblem('hidden width > 128 is not supported by the fused path')
    if widths[0] > 4 or widths[-1] > 8:
        raise UnsupportedProblem('at most 4 inputs and 8 outputs')
    # trainable scalar coefficients registered on the net (tedeous/models.py:183-195)
    own = {id(p) for l in linears for p in (l.weight, l.bias)}
    extra = [p for p in model.parameters() if id(p) not in own]
    return NetSpec(widths, linears, extra)


# ----------------------------------------------------------------------------------------------------
# operators -> term tables
# ----------------------------------------------------------------------------------------------------
@dataclass
class FactorIR:
    var: int
    axes: Tuple[int, ...]        # () = value, (0, 0) = d2/dx0^2, (0, 1) = mixed partial (lowered by `lower_mixed`)
    pow: float
    dirvec: Optional[Tuple[float, ...]] = None   # set by `lower_mixed`: derivative of order len(axes) ALONG this vector


@dataclass
class TermIR:
    coeff: object                # float | torch.Tensor (per row) | nn.Parameter
    factors: List[FactorIR]
    coeff_fn: object = None      # (callable, rows) the per-row buffer was evaluated from (refresh_coeffs)
    coeff_slice: object = None   # (lo, hi) rows kept by this rank
    host_pows: object = None     # a callable 'pow' somewhere in the term: [pow_j (number | callable)] per factor; the term
                                 # is then assembled by torch from per-point factor fields (`host_terms` of the IR)


def _check_coeff_callable(fn, label):
    """A callable coefficient is evaluated into a buffer (at lowering, and again by `refresh_coeffs`): no gradient can
    flow into state it closes over.  The reference re-evaluates it inside the autograd graph on every step
    (tedeous/derivative.py:41-42, 114-115), so a closure over trainable tensors would silently train differently."""
    cells = [c.cell_contents for c in (getattr(fn, '__closure__', None) or ()) if c is not None]
    cells += list(getattr(fn, '__defaults__', None) or ())
    if getattr(fn, '__self__', None) is not None:
        cells.append(fn.__self__)
    for obj in cells:
        trainable = (isinstance(obj, torch.Tensor) and obj.requires_grad) or \
                    (isinstance(obj, torch.nn.Module) and any(p.requires_grad for p in obj.parameters()))
        if trainable:
            raise UnsupportedProblem(f'term {label!r}: the callable coefficient closes over trainable state; pass the '
                                     f'trainable quantity as an nn.Parameter coefficient (parameter_registr) instead')


def _pure_axes(spec) -> Tuple[int, ...]:
    if spec == [None] or spec is None or spec == []:
        return ()
    axes = tuple(a for a in spec if a is not None)
    return axes


def parse_operator(op: dict, rows: torch.Tensor, coeff_rows: Optional[Callable] = None,
                   allow_host: bool = False) -> List[TermIR]:
    """Unified operator dict -> list of TermIR.  `rows` are the points the operator is evaluated at
    (callable coefficients are evaluated on them once); `coeff_rows(tensor)` maps a full-grid tensor
    coefficient onto `rows`."""
    terms = []
    for label, term in op.items():
        dif_dir = list(term.keys())[1]
        specs, pows, vars_ = term[dif_dir], term['pow'], term['var']
        if not isinstance(specs, list) or (specs and not isinstance(specs[0], list)):
            specs = [specs]
        facs = []
        host = any(callable(pw) for pw in pows)
        if host and not allow_host:
            raise UnsupportedProblem(f"term {label!r}: callable 'pow' is supported in the equation operator only")
        for spec, pw, var in zip(specs, pows, vars_):
            if isinstance(var, (list, tuple)) and len(var) == 1:
                var = var[0]     # 'var': [0] next to a scalar 'pow' (example_weak_LotkaVolterra.py:59-64): equation_unify
                                 # wraps it once more and the reference indexes model(grid)[:, [0]] with the list
            facs.append(FactorIR(int(var), _pure_axes(spec), 1.0 if host else float(pw)))
        coeff = term['coeff']
        coeff_fn = None
        if isinstance(coeff, tuple):          # reference NN-prepared form (callable, grid)
            coeff = coeff[0]
        if isinstance(coeff, torch.nn.Parameter):
            pass
        elif callable(coeff):
            _check_coeff_callable(coeff, label)
            coeff_fn = (coeff, rows)
            coeff = coeff(rows).reshape(-1).detach()
        elif isinstance(coeff, torch.Tensor):
            coeff = coeff.reshape(-1).detach()
            if coeff.numel() == 1:
                coeff = float(coeff)
            elif coeff_rows is not None:
                coeff = coeff_rows(coeff)
        else:
            coeff = float(coeff)
        if isinstance(coeff, torch.Tensor) and not isinstance(coeff, torch.nn.Parameter) \
                and coeff.numel() != rows.shape[0]:
            raise ValueError(f'term {label!r}: coefficient has {coeff.numel()} entries for {rows.shape[0]} rows')
        terms.append(TermIR(coeff, facs, coeff_fn, host_pows=list(pows) if host else None))
    return terms


@dataclass
class JetSpec:
    """Jet directions.  A direction is an input axis (int: d^k/dx_a^k) or a vector (tuple of d floats: the k-th
    derivative ALONG that vector - mixed partials are linear combinations of those, see `lower_mixed`)."""
    dirs: List[Tuple[object, int]]       # (axis | vector, max order): axes first (sorted), then vectors

    @property
    def J(self):
        return 1 + sum(o for _, o in self.dirs)

    def channel(self, f) -> int:
        """Channel of a factor (FactorIR) or of a pure axes tuple."""
        axes = f if isinstance(f, tuple) else f.axes
        if not axes:
            return 0
        key = getattr(f, 'dirvec', None)
        if key is None:
            key = axes[0]
        base = 1
        for a, o in self.dirs:
            if a == key:
                return base + len(axes) - 1
            base += o
        raise KeyError(axes)

    def vector(self, i: int, d: int) -> List[float]:
        a = self.dirs[i][0]
        return [float(x) for x in a] if isinstance(a, tuple) else [1.0 if ax == a else 0.0 for ax in range(d)]


def jet_spec(term_lists: Sequence[List[TermIR]]) -> JetSpec:
    """Directions and orders the (already `lower_mixed`-ed) terms need."""
    order: Dict[object, int] = {}
    for terms in term_lists:
        for t in terms:
            for f in t.factors:
                if not f.axes:
                    continue
                if f.dirvec is None and len(set(f.axes)) != 1:
                    raise UnsupportedProblem(f'mixed partial derivative {list(f.axes)} reached the jet set unlowered')
                if len(f.axes) > MAX_ORDER:
                    raise UnsupportedProblem(f'derivative order {len(f.axes)} > {MAX_ORDER}')
                key = f.dirvec if f.dirvec is not None else f.axes[0]
                order[key] = max(order.get(key, 0), len(f.axes))
    dirs = sorted((k, o) for k, o in order.items() if not isinstance(k, tuple)) + \
        sorted((k, o) for k, o in order.items() if isinstance(k, tuple))
    js = JetSpec(dirs)
    if len(dirs) > MAX_DIRS or js.J > MAX_J:
        raise UnsupportedProblem(f'jet set too large (J = {js.J}, {len(dirs)} directions)')
    return js


def _mixed_plan(n: int, need_q: Sequence[int], have_a: int, have_b: int):
    """Mixed partials of total order n along two axes a, b: d^n / da^(n-q) db^q = sum_j c_j D_{v_j}^n with directions
    v = e_a, e_b or e_a + t e_b, because D_v^n = sum_k C(n, k) t^k d^n / da^(n-k) db^k.  Picks the cheapest set of
    directions (fewest added jet channels given the pure orders `have_a`, `have_b` the operator needs anyway) whose rows
    span every wanted unit vector e_q.  -> ([('a' | 'b' | t, ...)], {q: coefficients})."""
    from itertools import combinations
    from math import comb as binom
    pool = ['a', 'b', 1.0, -1.0, 2.0, -2.0, 0.5]

    def row(c):
        if c == 'a':
            return [1.0] + [0.0] * n
        if c == 'b':
            return [0.0] * n + [1.0]
        return [binom(n, k) * c ** k for k in range(n + 1)]

    def cost(c):
        return max(0, n - have_a) if c == 'a' else max(0, n - have_b) if c == 'b' else n
    best = None
    for size in range(1, len(pool) + 1):
        for sub in combinations(pool, size):
            A = np.array([row(c) for c in sub], dtype=np.float64)            # [dirs, n + 1]
            sol = {}
            for q in need_q:
                e = np.zeros(n + 1)
                e[q] = 1.0
                c, *_ = np.linalg.lstsq(A.T, e, rcond=None)
                if np.abs(A.T @ c - e).max() > 1e-9:
                    break
                sol[q] = c
            else:
                k = (sum(cost(c) for c in sub), size)
                if best is None or k < best[0]:
                    best = (k, list(sub), sol)
    if best is None:
        raise UnsupportedProblem(f'no direction set for mixed partials of order {n}')
    return best[1], best[2]


def lower_mixed(term_lists: Sequence[List[TermIR]], d: int) -> List[List[TermIR]]:
    """Rewrites every factor that is a MIXED partial (reference: any axis list, tedeous/derivative.py:92-97, e.g. [0, 1])
    as a linear combination of pure directional derivatives of the same order (polarisation), so that the kernels only
    ever propagate 1-D Taylor jets: a term `c * u_xy * rest` becomes `sum_j (c * c_j) * D_{v_j}^2 u * rest`.  A mixed
    factor under a power p = 2 or 3 is written as p factors of power 1 first, so the power of the sum comes out as the
    k^p products of its k directional derivatives (equal factors of a product are merged back into one power); other
    powers of a mixed partial are not a finite sum of channel products and raise.  Two axes per mixed partial."""
    pure: Dict[int, int] = {}
    mixed: Dict[Tuple[int, int, int], set] = {}
    for terms in term_lists:
        for t in terms:
            for f in t.factors:
                ax = sorted(set(f.axes))
                if f.dirvec is not None:
                    continue
                if len(ax) == 1:
                    pure[ax[0]] = max(pure.get(ax[0], 0), len(f.axes))
                elif len(ax) == 2:
                    if f.pow not in (1.0, 2.0, 3.0):
                        raise UnsupportedProblem(f'mixed partial {list(f.axes)} with power {f.pow} (1, 2 or 3 only)')
                    if len(f.axes) > MAX_ORDER:
                        raise UnsupportedProblem(f'derivative order {len(f.axes)} > {MAX_ORDER}')
                    mixed.setdefault((ax[0], ax[1], len(f.axes)), set()).add(sum(1 for x in f.axes if x == ax[1]))
                elif len(ax) > 2:
                    raise UnsupportedProblem(f'mixed partial {list(f.axes)} along more than two axes')
    if not mixed:
        return [list(terms) for terms in term_lists]
    plans = {}
    for (a, b, n), qs in sorted(mixed.items(), key=lambda kv: -kv[0][2]):      # highest order first: its directions
        cands, sol = _mixed_plan(n, sorted(qs), pure.get(a, 0), pure.get(b, 0))  # serve the lower orders for free
        vecs = []
        for c in cands:
            if c == 'a':
                pure[a] = max(pure.get(a, 0), n)
                vecs.append(a)
            elif c == 'b':
                pure[b] = max(pure.get(b, 0), n)
                vecs.append(b)
            else:
                v = [0.0] * d
                v[a], v[b] = 1.0, float(c)
                vecs.append(tuple(v))
        plans[(a, b, n)] = (vecs, sol)

    def expand(t: TermIR) -> List[TermIR]:
        for i, f in enumerate(t.factors):
            ax = sorted(set(f.axes))
            if len(ax) != 2 or f.dirvec is not None:
                continue
            if f.pow != 1.0:
                copies = [FactorIR(f.var, tuple(f.axes), 1.0) for _ in range(int(f.pow))]
                return expand(TermIR(t.coeff, t.factors[:i] + copies + t.factors[i + 1:], t.coeff_fn))
            n, q = len(f.axes), sum(1 for x in f.axes if x == ax[1])
            vecs, sol = plans[(ax[0], ax[1], n)]
            out = []
            for v, c in zip(vecs, sol[q]):
                if abs(c) < 1e-12:
                    continue
                nf = FactorIR(f.var, (v,) * n, 1.0) if isinstance(v, int) else FactorIR(f.var, tuple(f.axes), 1.0, dirvec=v)
                facs = t.factors[:i] + [nf] + t.factors[i + 1:]
                if isinstance(t.coeff, torch.nn.Parameter):
                    if abs(c - 1.0) > 1e-12:
                        raise UnsupportedProblem('mixed partial under a trainable coefficient')
                    nt = TermIR(t.coeff, facs)
                elif isinstance(t.coeff, torch.Tensor):
                    fn = None
                    if t.coeff_fn is not None:
                        fn = (lambda rows, g=t.coeff_fn[0], c=float(c): c * g(rows), t.coeff_fn[1])
                    nt = TermIR(t.coeff * float(c), facs, fn)
                else:
                    nt = TermIR(float(t.coeff) * float(c), facs)
                out += expand(nt)                     # further mixed factors of the same term
            return out
        return [t]

    def merged(t: TermIR) -> TermIR:
        facs: List[FactorIR] = []
        for f in t.factors:
            for g in facs:
                if (g.var, g.axes, g.dirvec) == (f.var, f.axes, f.dirvec) and f.dirvec is not None:
                    g.pow += f.pow
                    break
            else:
                facs.append(FactorIR(f.var, f.axes, f.pow, dirvec=f.dirvec) if f.dirvec is not None else f)
        t.factors = facs
        return t

    def lowered(t: TermIR) -> List[TermIR]:
        out = expand(t)
        return out if len(out) == 1 and out[0] is t else [merged(e) for e in out]
    return [[e for t in terms for e in lowered(t)] for terms in term_lists]


# ----------------------------------------------------------------------------------------------------
# segments
# ----------------------------------------------------------------------------------------------------
@dataclass
class SegmentIR:
    name: str
    points: torch.Tensor                 # [n_groups * K, d], group-major
    K: int
    jet: JetSpec
    comb: Optional[np.ndarray]           # [M, K * J] or None (identity)
    chan_of: Callable[[FactorIR], int]   # factor -> virtual channel index
    cols: List[List[TermIR]]             # one term list per residual column
    slots: List[int]
    targets: Optional[torch.Tensor]      # [n_groups, n_cols]
    row_index: Optional[torch.Tensor] = None   # position of each group inside its field column
    n_groups_global: Optional[int] = None

    @property
    def n_groups(self):
        return self.points.shape[0] // self.K

    @property
    def M(self):
        return self.jet.J if self.comb is None else self.comb.shape[0]


def points_per_tile(J: int, K: int) -> int:
    """Points a 128-row tile holds: the largest multiple of lcm(4, K) with J * P <= 128."""
    pmax = ROWS_PER_TILE // J
    step = K * 4 // np.gcd(K, 4)
    p = (pmax // step) * step
    if p == 0:
        raise UnsupportedProblem(f'group of {K} points with {J} jet channels does not fit a tile')
    return int(p)


@dataclass
class ProblemIR:
    mode: str
    d: int
    net: NetSpec
    segments: List[SegmentIR]
    n_eq: int
    bnd_types: List[str]
    slot_len: List[int]                  # denominator of every slot (global)
    slot_lambda: List[float]
    n_interior: int

    @property
    def n_slots(self):
        return self.n_eq + len(self.bnd_types)

    def slot_scale(self) -> np.ndarray:
        return np.array([l / n for l, n in zip(self.slot_lambda, self.slot_len)], dtype=np.float64)


def _lambda_list(lam, n, what) -> List[float]:
    if isinstance(lam, torch.Tensor):
        lam = lam.reshape(-1).tolist()
    if isinstance(lam, (int, float)):
        return [float(lam)] * n
    lam = [float(x) for x in lam]
    if len(lam) != n:
        raise ValueError(f'{what}: expected {n} lambdas, got {len(lam)}')
    return lam


def _value_term(var: int) -> List[TermIR]:
    return [TermIR(1.0, [FactorIR(int(var), (), 1.0)])]


def _is_linear(terms: List[TermIR]) -> bool:
    return all(len(t.factors) == 1 and t.factors[0].pow == 1.0 and not isinstance(t.coeff, torch.nn.Parameter)
               for t in terms)


def _stencil_segment(name, bnd, terms, target, slot, h, variant, type_name, row_index) -> SegmentIR:
    """NN-mode boundary operator on points of one point type: literal one-sided shifted evaluations
    (tedeous/finite_diffs.py:120-223 via input_preprocessing.py:371-408, eval.py:264-281)."""
    d = bnd.shape[1]
    # distinct derivative specs -> virtual channels; distinct shift vectors -> group points
    chan_specs: List[Tuple[int, ...]] = [()]
    for t in terms:
        for f in t.factors:
            if f.axes not in chan_specs:
                chan_specs.append(f.axes)
    shift_list: List[Tuple[int, ...]] = []
    weights: List[Dict[int, float]] = []
    for axes in chan_specs:
        if not axes:
            sch, sg = [[0] * d], [1.0]
        else:
            sch, sg = Finite_diffs(list(axes), d, type_name).scheme_choose(variant, h=h)
        w: Dict[int, float] = {}
        for s, c in zip(sch, sg):
            s = tuple(s)
            if s not in shift_list:
                shift_list.append(s)
            k = shift_list.index(s)
            w[k] = w.get(k, 0.0) + float(c)
        weights.append(w)
    K, M = len(shift_list), len(chan_specs)
    if K > MAX_K or M > MAX_M:
        raise UnsupportedProblem(f'{name}: stencil too large (K = {K}, M = {M})')
    comb = np.zeros((M, K), dtype=np.float64)
    for m, w in enumerate(weights):
        for k, c in w.items():
            comb[m, k] = c
    # group-major points: x + shift * h, computed like Points_type.shift_points (tedeous/points_type.py:22-37)
    pts = bnd.unsqueeze(1).repeat(1, K, 1)
    for k, s in enumerate(shift_list):
        for a, mult in enumerate(s):
            if mult != 0:
                pts[:, k, a] = bnd[:, a] + mult * h
    return SegmentIR(name, pts.reshape(-1, d).contiguous(), K, JetSpec([]), comb,
                     lambda f: chan_specs.index(f.axes), [terms], [slot], target.reshape(-1, 1),
                     row_index=row_index)


def _interior_literal_fd(name, pts, cols, slots, h) -> SegmentIR:
    """Optional literal NN-mode interior evaluation: central differences with step h as shifted forwards
    (tedeous/derivative.py:30-58) instead of exact jets."""
    d = pts.shape[1]
    chan_specs: List[Tuple[int, ...]] = [()]
    for terms in cols:
        for t in terms:
            for f in t.factors:
                if f.axes not in chan_specs:
                    chan_specs.append(f.axes)
    shift_list, weights = [], []
    for axes in chan_specs:
        sch, sg = ([[0] * d], [1.0]) if not axes else Finite_diffs(list(axes), d, 'central').scheme_choose('1', h=h)
        w = {}
        for s, c in zip(sch, sg):
            s = tuple(s)
            if s not in shift_list:
                shift_list.append(s)
            k = shift_list.index(s)
            w[k] = w.get(k, 0.0) + float(c)
        weights.append(w)
    K, M = len(shift_list), len(chan_specs)
    if K > MAX_K or M > MAX_M:
        raise UnsupportedProblem(f'{name}: stencil too large (K = {K}, M = {M})')
    comb = np.zeros((M, K))
    for m, w in enumerate(weights):
        for k, c in w.items():
            comb[m, k] = c
    p = pts.unsqueeze(1).repeat(1, K, 1)
    for k, s in enumerate(shift_list):
        for a, mult in enumerate(s):
            if mult != 0:
                p[:, k, a] = pts[:, a] + mult * h
    return SegmentIR(name, p.reshape(-1, d).contiguous(), K, JetSpec([]), comb,
                     lambda f: chan_specs.index(f.axes), cols, slots, None)


def lower_problem(mode: str, grid: torch.Tensor, prepared_operator: List[dict], bconds: Optional[List[dict]],
                  model: torch.nn.Module, lambda_operator, lambda_bound, h: float = 0.001,
                  inner_order: str = '1', boundary_order: str = '2', nn_interior: str = 'jet',
                  shard: Tuple[int, int] = (0, 1)) -> ProblemIR:
    """Build the IR for modes 'NN' and 'autograd'.

    `shard = (rank, world)` keeps only this rank's contiguous block of every segment's rows; slot
    denominators stay global so per-rank partial losses / gradients simply add."""
    if mode not in ('NN', 'autograd'):
        raise ValueError(mode)
    net = net_spec(model)
    d = grid.shape[1]
    if d != net.widths[0]:
        raise ValueError(f'grid has {d} columns but the network takes {net.widths[0]} inputs')
    n_eq = len(prepared_operator)
    if n_eq > MAX_COLS:
        raise UnsupportedProblem(f'at most {MAX_COLS} equations')
    segments: List[SegmentIR] = []

    # ---- interior residual --------------------------------------------------------------------------
    if mode == 'NN':
        central = Points_type(grid).central_mask()
        pts = grid[central].contiguous()
        coeff_rows = (lambda c: c[central] if c.numel() == grid.shape[0] else c)
    else:
        central = None
        pts = grid.contiguous()
        coeff_rows = None
    cols = [parse_operator(eq, pts, coeff_rows, allow_host=True) for eq in prepared_operator]
    n_interior = pts.shape[0]
    # A callable 'pow' is an arbitrary torch function of the running product (tedeous/derivative.py:52-55, 126-129; shipped:
    # examples/examples_heat/example_heat_2d_long_time.py:94-99): such terms leave the term table.  The kernel evaluates
    # their factors as extra per-point field columns, torch applies the callables (autograd), and the parameter gradient
    # comes back through the kernel in vector-Jacobian mode (Solution._evaluate_hybrid).
    host_terms, extra_cols, extra_key = [], [], {}
    for e, terms in enumerate(cols):
        keep = []
        for t in terms:
            if t.host_pows is None:
                keep.append(t)
                continue
            chain = []
            for f, pw in zip(t.factors, t.host_pows):
                key = (f.var, f.axes)
                if key not in extra_key:
                    extra_key[key] = n_eq + len(extra_cols)
                    extra_cols.append([TermIR(1.0, [FactorIR(f.var, f.axes, 1.0)])])
                chain.append((extra_key[key], pw))
            host_terms.append((e, t.coeff, chain))
        cols[e] = keep
    if host_terms:
        if nn_interior == 'literal' and mode == 'NN':
            raise UnsupportedProblem("callable 'pow' with nn_interior='literal'")
        if shard[1] > 1:
            raise UnsupportedProblem("callable 'pow' sharded over ranks (the loss is assembled from per-point fields)")
        if n_eq + len(extra_cols) > MAX_COLS:
            raise UnsupportedProblem(f"callable 'pow': {n_eq} equations + {len(extra_cols)} factor fields > {MAX_COLS} columns")
        cols = cols + extra_cols
    if mode == 'NN' and nn_interior == 'literal':
        seg = _interior_literal_fd('operator', pts, cols, list(range(n_eq)), h)
    else:
        cols = lower_mixed(cols, d)
        js = jet_spec(cols)
        seg = SegmentIR('operator', pts, 1, js, None, lambda f, js=js: js.channel(f), cols,
                        list(range(n_eq)) + [0] * len(extra_cols), None)
    segments.append(seg)

    # ---- boundary rows ------------------------------------------------------------------------------
    bnd_types: List[str] = []
    type_len: Dict[str, int] = {}
    if not bconds:
        raise UnsupportedProblem('a problem without boundary conditions has no finite loss in the reference '
                                 '(losses.py:250-255)')
    ptype = Points_type(grid) if mode == 'NN' else None
    for ci, bc in enumerate(bconds):
        kind = bc['type']
        if kind not in bnd_types:
            bnd_types.append(kind)
            type_len[kind] = 0
        slot = n_eq + bnd_types.index(kind)
        bop, var = bc['bop'], bc['var']
        name = f'bc{ci}:{kind}'
        if kind == 'periodic':
            sides = [b.to(grid.dtype) for b in bc['bnd']]
            n = sides[0].shape[0]
            if any(s.shape[0] != n for s in sides):
                raise ValueError(f'{name}: periodic sides must have the same number of points')
            target = torch.zeros(n, dtype=grid.dtype, device=grid.device)   # bval [0.] zero-pads (q2)
            K = len(sides)
            sign = [1.0] + [-1.0] * (K - 1)
            pts_g = torch.stack(sides, dim=1).reshape(-1, d).contiguous()
            if bop is None:
                comb = np.array([sign])
                segments.append(SegmentIR(name, pts_g, K, JetSpec([]), comb, lambda f: 0,
                                          [_value_term(var)], [slot], target.reshape(-1, 1),
                                          row_index=torch.arange(type_len[kind], type_len[kind] + n)))
            else:
                if mode == 'NN':
                    raise UnsupportedProblem(f'{name}: NN-mode periodic condition with an operator')
                terms = parse_operator(bop, sides[0])
                if not _is_linear(terms) or any(isinstance(t.coeff, torch.Tensor) for t in terms):
                    raise UnsupportedProblem(f'{name}: periodic operator must be linear with constant coefficients')
                terms, = lower_mixed([terms], d)
                js = jet_spec([terms])
                J = js.J
                comb = np.zeros((J, K * J))
                for k in range(K):
                    for c in range(J):
                        comb[c, k * J + c] = sign[k]
                segments.append(SegmentIR(name, pts_g, K, js, comb, lambda f, js=js: js.channel(f),
                                          [terms], [slot], target.reshape(-1, 1),
                                          row_index=torch.arange(type_len[kind], type_len[kind] + n)))
            type_len[kind] += n
            continue

        bnd = bc['bnd'].to(grid.dtype)
        n = bnd.shape[0]
        target = bc['bval'].reshape(-1).to(grid.dtype)
        if target.numel() != n:
            raise ValueError(f'{name}: {target.numel()} target values for {n} boundary points')
        base = type_len[kind]
        if kind == 'dirichlet' or (kind == 'data' and bop is None):
            segments.append(SegmentIR(name, bnd.contiguous(), 1, JetSpec([]), None, lambda f: 0,
                                      [_value_term(var)], [slot], target.reshape(-1, 1),
                                      row_index=torch.arange(base, base + n)))
        elif kind in ('operator', 'data', 'robin'):
            if kind == 'robin':
                if mode == 'NN':
                    raise UnsupportedProblem('NN-mode robin conditions fail in the reference too (eval.py:357-388)')
                terms = _robin_terms(bop, bnd, var)
            else:
                terms = None
            if mode == 'autograd':
                terms, = lower_mixed([terms or parse_operator(bop, bnd)], d)
                js = jet_spec([terms])
                segments.append(SegmentIR(name, bnd.contiguous(), 1, js, None,
                                          lambda f, js=js: js.channel(f), [terms], [slot],
                                          target.reshape(-1, 1), row_index=torch.arange(base, base + n)))
            else:
                _, names = ptype.bnd_types(bnd)
                order = []
                for nm in names:
                    if nm not in order:
                        order.append(nm)
                for nm in order:
                    idx = torch.tensor([i for i, x in enumerate(names) if x == nm], device=bnd.device)
                    sub = bnd[idx]
                    terms_t = parse_operator(bop, sub, lambda c, idx=idx: c[idx] if c.numel() == n else c)
                    variant = inner_order if nm == 'central' else boundary_order
                    segments.append(_stencil_segment(f'{name}:{nm}', sub, terms_t, target[idx], slot, h,
                                                     variant, nm, base + idx.cpu()))
        else:
            raise ValueError(f'unknown condition type {kind!r}')
        type_len[kind] += n

    max_len = max(type_len.values())
    lam_op = _lambda_list(lambda_operator, n_eq, 'lambda_operator')
    lam_b = _lambda_list(lambda_bound, len(bnd_types), 'lambda_bound')
    ir = ProblemIR(mode, d, net, segments, n_eq, bnd_types,
                   [n_interior] * n_eq + [max_len] * len(bnd_types), lam_op + lam_b, n_interior)
    ir.type_len = [type_len[t] for t in bnd_types]
    # time slices of the causal loss: unique values of column 0 of the interior rows (tedeous/solution.py:51-57)
    ir.n_t = int(torch.unique(pts[:, 0]).numel())
    ir.interior_points = pts             # rows of the operator residual (NN mode: the central points), weak form
    ir.host_terms = host_terms           # [(equation, coeff, [(field column, pow | callable)])]: callable-'pow' terms
    for s in ir.segments:
        s.n_groups_global = s.n_groups
    if shard[1] > 1:
        _shard_segments(ir, *shard)
    return ir


def _robin_terms(bop: dict, bnd: torch.Tensor, var: int) -> List[TermIR]:
    """value = alpha * u + sum_beta beta * (whole bop applied) with alpha, betas = the terms' coefficients
    (tedeous/eval.py:357-388; the alpha*u term is counted again inside every beta term - SURVEY B.1 q5)."""
    coeffs = [bop[k]['coeff'] for k in bop]
    alpha, betas = coeffs[0], coeffs[1:]
    base = parse_operator(bop, bnd)

    def scale(c, t):
        c = c(bnd).reshape(-1) if callable(c) else c
        if isinstance(t.coeff, torch.nn.Parameter):
            raise UnsupportedProblem('robin condition with trainable coefficients')
        return t.coeff * c

    a = alpha(bnd).reshape(-1) if callable(alpha) else float(alpha)
    terms = [TermIR(a, [FactorIR(int(var), (), 1.0)])]
    for beta in betas:
        for t in base:
            terms.append(TermIR(scale(beta, t), list(t.factors)))
    return terms


def _shard_segments(ir: ProblemIR, rank: int, world: int) -> None:
    for s in ir.segments:
        n = s.n_groups
        lo, hi = (n * rank) // world, (n * (rank + 1)) // world
        s.points = s.points[lo * s.K: hi * s.K].contiguous()
        if s.targets is not None:
            s.targets = s.targets[lo:hi].contiguous()
        if s.row_index is not None:
            s.row_index = s.row_index[lo:hi]
        for terms in s.cols:
            for t in terms:
                if isinstance(t.coeff, torch.Tensor) and not isinstance(t.coeff, torch.nn.Parameter):
                    t.coeff = t.coeff[lo:hi].contiguous()
                    t.coeff_slice = (lo, hi)
        s.shard_range = (lo, hi)


# ----------------------------------------------------------------------------------------------------
# flattening for the C ABI
# ----------------------------------------------------------------------------------------------------
@dataclass
class FlatIR:
    seg: np.ndarray
    terms: np.ndarray
    factors: np.ndarray
    comb: np.ndarray                 # float32
    points: torch.Tensor             # [total_points, d] float32 on the device
    targets: torch.Tensor            # float32
    coeffs: torch.Tensor             # float32
    n_fields: int
    cparam_index: Dict[int, int]
    coeff_fns: list = None           # [(offset, numel, callable, rows, (lo, hi) | None, segment)]: callable-coefficient buffers
    inputs: torch.Tensor = None      # points | targets | coeffs as one buffer (the three tensors above are views of it)


def flatten(ir: ProblemIR, device) -> FlatIR:
    seg = np.zeros(len(ir.segments), dtype=SEGMENT_DTYPE)
    terms, factors, comb = [], [], []
    pts, tgts, coefs = [], [], []
    pts_off = tgt_off = coef_off = field_off = 0
    coeff_fns = []
    cparam_index = {id(p): i for i, p in enumerate(ir.net.coeff_params)}
    for si, s in enumerate(ir.segments):
        r = seg[si]
        J = s.jet.J
        r['n_groups'], r['pts_off'], r['K'], r['M'], r['n_cols'] = s.n_groups, pts_off, s.K, s.M, len(s.cols)
        r['n_dirs'] = len(s.jet.dirs)
        for i, (a, o) in enumerate(s.jet.dirs):
            r['dir_axis'][i], r['dir_order'][i] = (-1 if isinstance(a, tuple) else a), o
            r['dir_vec'][i][:ir.d] = s.jet.vector(i, ir.d)
        points_per_tile(J, s.K)                       # validates the tile shape
        if s.M > J * s.K:
            raise UnsupportedProblem(f'{s.name}: more virtual channels than evaluated jets')
        r['identity'] = 1 if s.comb is None else 0
        r['comb_off'] = sum(c.size for c in comb)
        if s.comb is not None:
            assert s.comb.shape == (s.M, s.K * J), (s.comb.shape, s.M, s.K, J)
            comb.append(np.asarray(s.comb, dtype=np.float32).reshape(-1))
        r['field_off'] = field_off
        field_off += s.n_groups * len(s.cols)
        if s.targets is not None:
            r['tgt_off'] = tgt_off
            tgts.append(s.targets.reshape(-1).to(device=device, dtype=torch.float32))
            tgt_off += s.targets.numel()
        else:
            r['tgt_off'] = -1
        for ci, (tl, slot) in enumerate(zip(s.cols, s.slots)):
            r['col_term_begin'][ci] = len(terms)
            for t in tl:
                fb = len(factors)
                for f in t.factors:
                    ip = int(f.pow) if float(f.pow).is_integer() and 0 <= f.pow <= 16 else -1
                    factors.append((f.var, s.chan_of(f), f.pow, ip))
                    if f.var >= ir.net.widths[-1]:
                        raise ValueError(f'{s.name}: var {f.var} but the network has {ir.net.widths[-1]} outputs')
                if isinstance(t.coeff, torch.nn.Parameter):
                    if id(t.coeff) not in cparam_index:
                        raise UnsupportedProblem('trainable coefficient is not registered on the network '
                                                 '(use parameter_registr)')
                    terms.append((0.0, COEFF_PARAM, cparam_index[id(t.coeff)], fb, len(factors)))
                elif isinstance(t.coeff, torch.Tensor):
                    terms.append((0.0, COEFF_BUFFER, coef_off, fb, len(factors)))
                    if t.coeff_fn is not None:
                        coeff_fns.append((coef_off, t.coeff.numel(), t.coeff_fn[0], t.coeff_fn[1], t.coeff_slice, si))
                    coefs.append(t.coeff.reshape(-1).to(device=device, dtype=torch.float32))
                    coef_off += t.coeff.numel()
                else:
                    terms.append((float(t.coeff), COEFF_CONST, 0, fb, len(factors)))
            r['col_term_end'][ci] = len(terms)
            r['col_slot'][ci] = slot
        pts.append(s.points.to(device=device, dtype=torch.float32))
        pts_off += s.points.shape[0]

    def cat(lst):
        return torch.cat(lst).contiguous() if lst else torch.zeros(1, dtype=torch.float32, device=device)
    # the per-step inputs - points, targets, coefficient buffers - are views of ONE device buffer (`inputs`), so a host
    # that refreshes them every step (a data loader, bench.py's end-to-end leg) needs a single host-to-device copy
    p_all, t_all, c_all = torch.cat(pts).contiguous(), cat(tgts), cat(coefs)
    d_in = p_all.shape[1]
    n_p, n_t, n_c = p_all.numel(), t_all.numel(), c_all.numel()
    pad = lambda n: (n + 3) // 4 * 4
    inputs = torch.zeros(pad(n_p) + pad(n_t) + pad(n_c), dtype=torch.float32, device=device)
    inputs[:n_p] = p_all.reshape(-1)
    inputs[pad(n_p):pad(n_p) + n_t] = t_all
    inputs[pad(n_p) + pad(n_t):pad(n_p) + pad(n_t) + n_c] = c_all
    p_all = inputs[:n_p].view(-1, d_in)
    t_all = inputs[pad(n_p):pad(n_p) + n_t]
    c_all = inputs[pad(n_p) + pad(n_t):pad(n_p) + pad(n_t) + n_c]
    flat = FlatIR(seg, np.array(terms, dtype=TERM_DTYPE) if terms else np.zeros(0, TERM_DTYPE),
                  np.array(factors, dtype=FACTOR_DTYPE) if factors else np.zeros(0, FACTOR_DTYPE),
                  np.concatenate(comb).astype(np.float32) if comb else np.zeros(1, np.float32),
                  p_all, t_all, c_all, field_off, cparam_index, coeff_fns)
    flat.inputs = inputs
    return flat
