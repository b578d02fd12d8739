"""Host logic: the IR produced by plan.lower_problem, executed in fp64 on the CPU (tests/ir_interp.py), must
reproduce the reference's loss / gradient.  autograd-mode problems agree to rounding; NN-mode problems
agree to the O(h^2) truncation of the central differences the jets replace (boundary stencils are literal)."""
import numpy as np
import pytest
import torch

import problems
from helpers import build, load_golden
from ir_interp import evaluate_ir
from torch_de_solver_b200.input_preprocessing import Operator_bcond_preproc
from torch_de_solver_b200.plan import lower_problem, flatten, points_per_tile
from torch_de_solver_b200.solution import deepcopy_equation

NET_CASES = sorted(k for k in problems.ZOO if 'mat' not in k and 'weak' not in k)
WEAK_CASES = sorted(k for k in problems.ZOO if 'weak' in k)


def lower(name, dtype='float64', nn_interior='jet', shard=(0, 1), weights=None):
    g = load_golden(name, dtype)
    prob, grid, bconds, model = build(name, dtype, g['weights'] if weights is None else weights)
    kw = prob.compile_kwargs
    eq = Operator_bcond_preproc(grid, prob.equation.equation_lst, bconds, h=kw.get('h', 0.001)).set_strategy(prob.mode)
    eq = deepcopy_equation(eq)
    ir = lower_problem(prob.mode, grid, eq.operator_prepare(), eq.bnd_prepare(), model, kw['lambda_operator'],
                       kw['lambda_bound'], h=kw.get('h', 0.001), nn_interior=nn_interior, shard=shard)
    return g, prob, model, ir


@pytest.mark.parametrize('name', NET_CASES)
def test_ir_matches_reference(name):
    g, prob, model, ir = lower(name)
    tol = prob.compile_kwargs.get('tol', 0)
    loss, loss_n, mse, _ = evaluate_ir(ir, model, tol=tol)
    params = list(model.parameters())
    grads = torch.autograd.grad(loss, params)
    grad = torch.cat([x.reshape(-1) for x in grads]).numpy()
    exact = prob.mode == 'autograd'
    rel = 1e-10 if exact else 2e-4          # NN: central-difference truncation (h^2 u''' / 6 ...), see DESIGN.md
    assert float(loss) == pytest.approx(float(g['loss']), rel=rel)
    assert float(loss_n) == pytest.approx(float(g['loss_normalized']), rel=rel)
    n_eq = ir.n_eq
    if tol == 0:                            # causal: the slot sums are weighted, the fixture's op_mse is not
        np.testing.assert_allclose([float(m) for m in mse[:n_eq]], g['op_mse'], rtol=1e-9 if exact else 2e-3)
    # boundary columns: the reference's mean is over the padded length = max over types
    np.testing.assert_allclose([float(m) for m in mse[n_eq:]], g['bval_mse'], rtol=1e-9 if exact else 1e-9)
    gn = np.linalg.norm(g['grad'])
    assert np.linalg.norm(grad - g['grad']) <= (1e-9 if exact else 5e-4) * gn
    assert ir.bnd_types == [str(k) for k in g['bval_keys']]
    assert ir.type_len == [int(x) for x in g['bval_length']]


@pytest.mark.parametrize('name', [k for k in NET_CASES if '_NN' in k and 'mix' not in k and 'callable' not in k])
def test_literal_fd_interior_matches_reference_fp64(name):
    """nn_interior='literal' restates NN mode as shifted evaluations: agrees with the fp64 reference to rounding."""
    g, prob, model, ir = lower(name, nn_interior='literal')
    tol = prob.compile_kwargs.get('tol', 0)
    loss, loss_n, mse, _ = evaluate_ir(ir, model, tol=tol)
    assert float(loss) == pytest.approx(float(g['loss']), rel=1e-9)
    if tol == 0:
        np.testing.assert_allclose([float(m) for m in mse[:ir.n_eq]], g['op_mse'], rtol=1e-7)


def test_sharded_ir_sums_to_full():
    name = 'nonlinear_mix_autograd'
    g, prob, model, ir = lower(name)
    full, _, _, _ = evaluate_ir(ir, model)
    parts = []
    for r in range(3):
        _, _, _, ir_r = lower(name, shard=(r, 3))
        parts.append(float(evaluate_ir(ir_r, model)[0]))
    assert sum(parts) == pytest.approx(float(full), rel=1e-12)


def test_flatten_layout():
    g, prob, model, ir = lower('navier_stokes_autograd')
    flat = flatten(ir, torch.device('cpu'))
    assert len(flat.seg) == len(ir.segments)
    seg0 = flat.seg[0]
    assert seg0['identity'] == 1 and seg0['K'] == 1 and seg0['n_cols'] == 3
    assert list(seg0['dir_axis'][:3]) == [0, 1, 2] and list(seg0['dir_order'][:3]) == [2, 2, 1]
    assert flat.points.shape[0] == sum(s.points.shape[0] for s in ir.segments)
    assert int(seg0['n_groups']) == 729
    # forcing term is a per-row buffer
    assert (flat.terms['kind'] == 1).sum() == 1
    assert points_per_tile(6, 1) == 20 and points_per_tile(4, 1) == 32 and points_per_tile(1, 3) == 120


@pytest.mark.parametrize('name', WEAK_CASES)
def test_weak_form_matches_reference(name):
    """Weak-form loss (losses.py:184-228): the per-point fields of the lowered IR, the product's vectorised nested
    integration (torch_de_solver_b200.losses) and the boundary MSE reproduce the reference's loss and gradient."""
    from torch_de_solver_b200.losses import Losses, weak_operator
    g, prob, model, ir = lower(name)
    _, _, _, fields = evaluate_ir(ir, model)
    kw = prob.compile_kwargs
    wop = weak_operator(fields[0], ir.interior_points, kw['weak_form'])
    n_types, max_len = len(ir.bnd_types), max(ir.type_len)
    bval = torch.zeros(max_len, n_types, dtype=torch.float64)
    tval = torch.zeros_like(bval)
    for s, f in zip(ir.segments[1:], fields[1:]):
        col = s.slots[0] - ir.n_eq
        bval = bval.index_put((s.row_index, torch.full_like(s.row_index, col)), f[:, 0])
        tval[s.row_index, col] = s.targets.reshape(-1).double()
    loss, loss_n = Losses(prob.mode, kw['weak_form'], None, 0).compute(wop, bval, tval, kw['lambda_operator'], kw['lambda_bound'])
    grads = torch.autograd.grad(loss.sum(), list(model.parameters()))
    grad = torch.cat([x.reshape(-1) for x in grads]).numpy()
    exact = prob.mode == 'autograd'
    assert tuple(loss.shape) == (1, 1)                     # the reference's shape for this loss
    assert float(loss) == pytest.approx(float(g['loss']), rel=1e-10 if exact else 2e-4)
    assert float(loss_n) == pytest.approx(float(g['loss_normalized']), rel=1e-10 if exact else 2e-4)
    np.testing.assert_allclose(wop.detach().reshape(-1).numpy(), g['op_mse'], rtol=1e-9 if exact else 2e-3)
    gn = np.linalg.norm(g['grad'])
    assert np.linalg.norm(grad - g['grad']) <= (1e-9 if exact else 5e-4) * gn


@pytest.mark.parametrize('d', [1, 2, 3])
def test_vectorised_integration_equals_the_reference_loop(d):
    """torch_de_solver_b200.losses.integration against the loop-for-loop restatement in the oracle (eval.py:13-52),
    including runs of one row and a ragged run structure."""
    from oracle import tedeous_oracle as orc
    from torch_de_solver_b200.losses import integration
    torch.manual_seed(d)
    axes = [torch.linspace(0, 1, n, dtype=torch.float64) ** 1.3 for n in (5, 4, 6)[:d]]
    grid = torch.cartesian_prod(*axes).reshape(-1, d)
    if d > 1:
        keep = torch.ones(len(grid), dtype=torch.bool)
        keep[3] = keep[7] = False                          # ragged runs
        keep[-6:-1] = False                                # a run of a single row
        grid = grid[keep]
    f = torch.randn(len(grid), dtype=torch.float64)
    want, wgrid = orc.integration(f, grid)
    got, ggrid = integration(f, grid)
    if d == 1:
        assert float(got) == pytest.approx(float(want), rel=1e-12)
    else:
        np.testing.assert_allclose(got.numpy(), np.array([float(x) for x in want]), rtol=1e-12, atol=1e-15)
        assert torch.equal(ggrid, wgrid)


def test_scheme_choose_equals_reference_table():
    """Finite_diffs.scheme_choose against the committed output of the unmodified reference
    (tests/golden/make_fd_table.py -> finite_diffs_table.json; tedeous/finite_diffs.py:244-268): identical shift
    lists in the same order, weights equal to the last bit."""
    import json
    import os
    from torch_de_solver_b200.finite_diffs import Finite_diffs
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'finite_diffs_table.json')
    table = json.load(open(path))
    assert len(table) > 1000
    for c in table:
        got = Finite_diffs(c['term'], c['nvars'], c['type']).scheme_choose(c['order'], h=c['h'])
        assert got[0] == c['scheme'], c
        assert got[1] == c['sign'], c


@pytest.mark.parametrize('n,qs,have', [(2, [1], (2, 2)), (2, [1], (0, 0)), (3, [1], (2, 1)), (3, [1, 2], (3, 3)),
                                       (4, [1, 2, 3], (4, 4)), (4, [2], (2, 2))])
def test_mixed_partials_as_directional_derivatives(n, qs, have):
    """`plan._mixed_plan`: d^n / dx^(n-q) dy^q = sum_j c_j D_{v_j}^n (polarisation) - checked on a random polynomial of
    degree 5 in two variables with torch autograd in fp64, for every direction set the search returns."""
    import itertools
    from torch_de_solver_b200.plan import _mixed_plan
    torch.manual_seed(n * 10 + len(qs))
    coef = torch.randn(6, 6, dtype=torch.float64)

    def poly(p):
        x, y = p[..., 0], p[..., 1]
        return sum(coef[i, j] * x ** i * y ** j for i, j in itertools.product(range(6), range(6)) if i + j <= 5)

    def directional(p, v, order):
        p = p.clone().requires_grad_(True)
        f = poly(p)
        for _ in range(order):
            g, = torch.autograd.grad(f.sum(), p, create_graph=True)
            f = g @ torch.as_tensor(v, dtype=torch.float64)
        return f.detach()

    def partial(p, axes):
        p = p.clone().requires_grad_(True)
        f = poly(p)
        for a in axes:
            g, = torch.autograd.grad(f.sum(), p, create_graph=True)
            f = g[..., a]
        return f.detach()
    pts = torch.rand(7, 2, dtype=torch.float64)
    cands, sol = _mixed_plan(n, qs, *have)
    vecs = [(1.0, 0.0) if c == 'a' else (0.0, 1.0) if c == 'b' else (1.0, float(c)) for c in cands]
    for q in qs:
        got = sum(float(c) * directional(pts, v, n) for c, v in zip(sol[q], vecs))
        ref = partial(pts, [0] * (n - q) + [1] * q)
        assert torch.allclose(got, ref, rtol=1e-9, atol=1e-9), (n, q, cands)
    # pure directions the operator needs anyway are free: with both present a second-order mixed partial adds ONE direction
    if n == 2 and have == (2, 2):
        assert sorted(map(str, cands)) == sorted(['a', 'b', '1.0'])
    if n == 2 and have == (0, 0):
        assert sorted(cands) == [-1.0, 1.0]


def test_power_of_a_mixed_partial_expands_into_products():
    """(u_xy)^2 with the polarisation u_xy = sum_j c_j D_j^2 u becomes sum_jk c_j c_k D_j^2 u D_k^2 u; a fractional power of a
    mixed partial has no such expansion and raises (no silent fallback)."""
    from torch_de_solver_b200.plan import FactorIR, TermIR, lower_mixed, UnsupportedProblem
    pure = TermIR(1.0, [FactorIR(0, (0, 0), 1.0), FactorIR(0, (1, 1), 1.0)])
    sq = TermIR(-1.0, [FactorIR(0, (0, 1), 2.0)])
    (out,) = lower_mixed([[pure, sq]], 2)
    assert out[0] is pure
    terms = out[1:]
    assert len(terms) in (4, 9)                    # k = 2 or 3 directions for u_xy, squared
    for t in terms:
        assert all(f.dirvec is not None or len(set(f.axes)) == 1 for f in t.factors)
        assert sum(f.pow for f in t.factors) == 2.0
    with pytest.raises(UnsupportedProblem):
        lower_mixed([[TermIR(1.0, [FactorIR(0, (0, 1), 1.5)])]], 2)
