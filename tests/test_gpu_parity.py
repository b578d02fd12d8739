"""GPU parity: the fused CUDA path (through Model.compile / Solution.evaluate -> the C ABI) against the golden
fixtures of the unmodified reference and against the oracle.

Tolerances are the north star's: loss <= 1e-5 relative, gradient norm <= 1e-4 relative; we also bound the
gradient *vector* error.  The reference values are the fp64 goldens (NN-mode fp32 finite differences carry
their own cancellation noise)."""
import ctypes as C

import numpy as np
import pytest
import torch

import problems
import torch_de_solver_b200 as tdb
from helpers import load_golden, oracle_eval, set_weights

pytestmark = pytest.mark.gpu
NET_CASES = sorted(k for k in problems.ZOO if 'mat' not in k and 'weak' not in k)
WEAK_CASES = sorted(k for k in problems.ZOO if 'weak' in k)
LOSS_RTOL, GRADNORM_RTOL, GRADVEC_RTOL = 1e-5, 1e-4, 2e-4


@pytest.fixture()
def cuda_default():
    torch.set_default_device('cuda:0')
    yield torch.device('cuda:0')
    torch.set_default_device('cpu')


def fused(name, weights, **opts):
    prob = problems.ZOO[name](tdb, 'float32')
    net = problems.make_net(prob.net_layers, torch.float32, prob.init)
    set_weights(list(net.parameters()), weights)
    net = net.to('cuda:0')
    model = tdb.Model(net, prob.domain, prob.equation, prob.conditions)
    model.compile(prob.mode, **prob.compile_kwargs, **opts)
    return prob, net, model.solution_cls


@pytest.mark.parametrize('name', NET_CASES)
def test_loss_and_gradient_match_reference(name, cuda_default):
    g = load_golden(name, 'float64')
    prob, net, sol = fused(name, g['weights'])
    loss, loss_n = sol.evaluate()
    assert loss.shape == (1,) and loss_n.shape == (1,)
    loss.backward()
    grad = torch.cat([p.grad.reshape(-1) for p in net.parameters()]).double().cpu().numpy()
    assert float(loss) == pytest.approx(float(g['loss']), rel=LOSS_RTOL)
    # NN mode: the interior residual uses exact derivative jets where the reference uses central differences of
    # step h, so the operator MSE differs by the O(h^2) truncation (2e-4 relative at h = 0.01, < 1e-6 at the
    # default h = 0.001); the lambda-weighted loss above stays inside the north-star tolerance.
    trunc = prob.mode == 'NN' and prob.compile_kwargs.get('h', 0.001) > 0.001
    assert float(loss_n) == pytest.approx(float(g['loss_normalized']), rel=3e-4 if trunc else LOSS_RTOL)
    gn = np.linalg.norm(g['grad'])
    assert abs(np.linalg.norm(grad) - gn) <= GRADNORM_RTOL * gn
    assert np.linalg.norm(grad - g['grad']) <= GRADVEC_RTOL * gn
    if not prob.compile_kwargs.get('tol', 0):      # causal loss: the slot sums are weighted, the fixture's are not
        np.testing.assert_allclose(sol.op_mse.cpu().numpy(), g['op_mse'], rtol=5e-4)
    # NN-mode one-sided boundary stencils are literal fp32 differences (3u - 4u + u) / 2h: cancellation noise
    np.testing.assert_allclose(sol.bval_mse.cpu().numpy(), g['bval_mse'], rtol=1e-3 if prob.mode == 'NN' else 1e-4)
    assert sol.bval_keys == [str(k) for k in g['bval_keys']]
    assert sol.bval_length == [int(x) for x in g['bval_length']]


@pytest.mark.parametrize('name', NET_CASES)
def test_fields_match_reference(name, cuda_default):
    """op / bval / true_bval as the callbacks read them (Solution attributes)."""
    g = load_golden(name, 'float64')
    prob, net, sol = fused(name, g['weights'])
    sol.evaluate()
    op = sol.op.cpu().double().numpy()
    assert op.shape[0] == int(g['op_rows'])
    scale = np.abs(g['op_head']).max() + 1e-12
    atol = 3e-3 * scale if prob.mode == 'NN' else 2e-4 * scale
    np.testing.assert_allclose(op[:256], g['op_head'], atol=atol, rtol=1e-3)
    bscale = np.abs(g['bval']).max() + 1e-12
    # targets are evaluated in fp32 on the device (golden: fp64)
    tscale = np.abs(g['true_bval']).max() + 1e-12
    np.testing.assert_allclose(sol.true_bval.cpu().numpy(), g['true_bval'], atol=2e-6 * tscale + 1e-7, rtol=1e-6)
    # NN-mode one-sided boundary stencils are literal fp32 differences divided by 2h: cancellation noise ~1e-4
    batol = 1e-3 if prob.mode == 'NN' else 1e-4
    np.testing.assert_allclose(sol.bval.cpu().numpy(), g['bval'], atol=batol * bscale, rtol=batol)


@pytest.mark.parametrize('name', ['burgers_NN_small', 'wave_NN', 'kdv_NN'])
def test_literal_fd_interior(name, cuda_default):
    """nn_interior='literal' (shifted evaluations, the reference's own arithmetic in fp32) agrees with the fp64
    reference up to fp32 cancellation noise of the finite differences."""
    g = load_golden(name, 'float64')
    prob, net, sol = fused(name, g['weights'], nn_interior='literal')
    loss, _ = sol.evaluate()
    assert float(loss) == pytest.approx(float(g['loss']), rel=2e-3)


# nets with more than two W x W layers keep Z / dW outside TMEM's 512 columns: they stay on the SIMT kernel
FOUR_DIRS = ('mixed_bbm_autograd',)     # 4 jet directions (x, t, x + t, x - t): served by the SIMT kernel only
TC_CASES = [k for k in NET_CASES if k not in ('navier_stokes_autograd', 'burgers_autograd_4h') + FOUR_DIRS]


@pytest.mark.parametrize('name', TC_CASES)
def test_tensor_core_path_matches_reference(name, cuda_default):
    """impl=2 forces the tcgen05 3xTF32 kernel for the interior segment (boundary segments stay on the SIMT kernel)."""
    g = load_golden(name, 'float64')
    prob, net, sol = fused(name, g['weights'], impl=2)
    assert sol._plan.launches_per_call >= 3
    loss, loss_n = sol.evaluate()
    loss.backward()
    grad = torch.cat([p.grad.reshape(-1) for p in net.parameters()]).double().cpu().numpy()
    assert float(loss) == pytest.approx(float(g['loss']), rel=LOSS_RTOL)
    gn = np.linalg.norm(g['grad'])
    assert abs(np.linalg.norm(grad) - gn) <= GRADNORM_RTOL * gn
    assert np.linalg.norm(grad - g['grad']) <= GRADVEC_RTOL * gn
    if not prob.compile_kwargs.get('tol', 0):
        np.testing.assert_allclose(sol.op_mse.cpu().numpy(), g['op_mse'], rtol=5e-4)
    # and against the SIMT fp32 kernel on the same inputs
    prob2, net2, sol2 = fused(name, g['weights'], impl=1)
    ref = sol2._run_plan()[0].double()            # (_run_plan refreshes the causal row weights when tol != 0)
    out = sol._run_plan()[0].double()
    assert float(out[0]) == pytest.approx(float(ref[0]), rel=2e-6)
    k = 2 + sol._n_slots
    # trained states: the gradient is a small difference of large per-point terms, and 3xTF32 (lo * lo dropped, lo
    # truncated to tf32) is ~7x noisier than fp32 FMAs there - still inside the north-star bounds checked above
    assert float((out[k:] - ref[k:]).norm()) <= (1.5e-4 if 'trained' in name else 2e-5) * float(ref[k:].norm())
    a, b = sol._plan.loss_grad(), sol._plan.loss_grad()
    assert torch.equal(a, b)                      # fixed accumulation order: bit-reproducible


@pytest.mark.parametrize('name', [k for k in NET_CASES if k not in FOUR_DIRS])
def test_streamed_tensor_core_path_matches_reference(name, cuda_default):
    """impl=3: jet_tcs_kernel (fused forward / backward-data on tcgen05, two tiles in flight, any depth) +
    wgrad_gemm_kernel (weight gradients of the W x W layers from the streamed Y / gZ rows)."""
    g = load_golden(name, 'float64')
    prob, net, sol = fused(name, g['weights'], impl=3)
    assert sol._plan.launches_per_call >= 4
    loss, loss_n = sol.evaluate()
    loss.backward()
    grad = torch.cat([p.grad.reshape(-1) for p in net.parameters()]).double().cpu().numpy()
    assert float(loss) == pytest.approx(float(g['loss']), rel=LOSS_RTOL)
    gn = np.linalg.norm(g['grad'])
    assert abs(np.linalg.norm(grad) - gn) <= GRADNORM_RTOL * gn
    assert np.linalg.norm(grad - g['grad']) <= GRADVEC_RTOL * gn
    if not prob.compile_kwargs.get('tol', 0):
        np.testing.assert_allclose(sol.op_mse.cpu().numpy(), g['op_mse'], rtol=5e-4)
    prob2, net2, sol2 = fused(name, g['weights'], impl=1)
    ref = sol2._run_plan()[0].double()
    out = sol._run_plan()[0].double()
    assert float(out[0]) == pytest.approx(float(ref[0]), rel=2e-6)
    k = 2 + sol._n_slots
    # trained states: the gradient is a small difference of large per-point terms, and 3xTF32 (lo * lo dropped, lo
    # truncated to tf32) is ~7x noisier than fp32 FMAs there - still inside the north-star bounds checked above
    assert float((out[k:] - ref[k:]).norm()) <= (1.5e-4 if 'trained' in name else 2e-5) * float(ref[k:].norm())
    a, b = sol._plan.loss_grad(), sol._plan.loss_grad()
    assert torch.equal(a, b)                      # fixed accumulation order: bit-reproducible


def test_auto_dispatch_reaches_tensor_cores(cuda_default):
    """impl = 0 (what Model.compile does by default) at the sizes the benchmark runs: from 4096 interior rows on the
    streamed tensor-core pair serves every eligible net (value-row boundary segments inside the interior launch); the
    numerics of these fixtures are checked by test_loss_and_gradient_match_reference (fp64 goldens of the reference)."""
    for name, lo in (('wave_autograd_1e5', 4), ('kdv_autograd_1e5', 4), ('wave_autograd_3e5', 4)):
        g = load_golden(name, 'float64')
        prob, net, sol = fused(name, g['weights'])
        assert sol._plan.launches_per_call >= lo, (name, sol._plan.launches_per_call)
        assert sol._plan.kernel_path.startswith('tcgen05-3xtf32 streamed'), sol._plan.kernel_path


def test_derivative_seam(cuda_default):
    """Derivative(model, p).set_strategy(mode).take_derivative(term, grid) (tedeous/derivative.py:326-363) against plain
    torch autograd in fp64 on the CPU; the lowered plan is cached per (term, points) and re-reads the live weights."""
    from torch_de_solver_b200.derivative import Derivative
    g = load_golden('wave_autograd', 'float64')
    prob, net, sol = fused('wave_autograd', g['weights'])
    grid = sol.grid
    net64 = problems.make_net(prob.net_layers, torch.float64, prob.init)
    set_weights(list(net64.parameters()), g['weights'])

    def reference(model, axes, coeff, power):
        pts = grid.detach().double().cpu().requires_grad_()
        fi = model(pts)[:, 0].sum()
        for ax in axes:
            grads, = torch.autograd.grad(fi, pts, create_graph=True)
            fi = grads[:, ax].sum()
        return (coeff * grads[:, axes[-1]] ** power).detach().reshape(-1, 1)

    for mode in ('autograd', 'NN'):
        strat = Derivative(net, 2).set_strategy(mode)
        term = {'coeff': 2.0, 'd2u/dx2': [[0, 0]], 'pow': [1], 'var': [0]}
        out = strat.take_derivative(term, grid)
        ref = reference(net64.cpu(), [0, 0], 2.0, 1)
        assert out.shape == ref.shape
        assert float((out.double().cpu() - ref).norm()) <= 2e-5 * float(ref.norm())
        assert len(strat._plans) == 1
        strat.take_derivative(term, grid)
        assert len(strat._plans) == 1                 # same term, same points: the cached plan
        term_t = {'coeff': 1.0, 'du/dt': [[1]], 'pow': [2], 'var': [0]}
        out_t = strat.take_derivative(term_t, grid)
        ref_t = reference(net64.cpu(), [1], 1.0, 2)
        assert float((out_t.double().cpu() - ref_t).norm()) <= 2e-5 * float(ref_t.norm())
        assert len(strat._plans) == 2
    # live weights: scale the last layer, the cached plan must see it
    with torch.no_grad():
        list(net.parameters())[-2].mul_(3.0)
    out3 = strat.take_derivative(term, grid)
    assert float((out3 - 3.0 * out).norm()) <= 1e-5 * float(out3.norm())


def test_callable_coefficients_refresh_and_trainable_closures(cuda_default):
    """Callable coefficients are evaluated into buffers: `refresh_coeffs()` / callable_coeffs='every_step' re-evaluate
    them (the reference calls them on every step, tedeous/derivative.py:41-42); closures over trainable tensors raise."""
    from torch_de_solver_b200.plan import UnsupportedProblem
    state = {'k': 1.0}

    def build(coeff, **opts):
        dom = tdb.Domain()
        dom.variable('x', [0, 1], 12, dtype='float32')
        dom.variable('t', [0, 1], 12, dtype='float32')
        bc = tdb.Conditions()
        bc.dirichlet({'x': [0, 1], 't': 0}, value=0.5)
        eq = tdb.Equation()
        eq.add({'du/dt': {'coeff': 1., 'du/dt': [1], 'pow': 1, 'var': 0},
                'k*u': {'coeff': coeff, 'u': [None], 'pow': 1, 'var': 0}})
        net = problems.make_net([2, 16, 16, 1], torch.float32).to('cuda:0')
        model = tdb.Model(net, dom, eq, bc)
        model.compile('autograd', lambda_operator=1, lambda_bound=10, **opts)
        return model.solution_cls

    sol = build(lambda g: state['k'] * g[:, 0])
    l1 = float(sol.evaluate()[0])
    state['k'] = 4.0
    assert float(sol.evaluate()[0]) == l1                       # 'once': the buffer still holds k = 1
    sol.refresh_coeffs()
    l4 = float(sol.evaluate()[0])
    assert abs(l4 - l1) > 1e-6 * abs(l1)
    state['k'] = 1.0
    sol_e = build(lambda g: state['k'] * g[:, 0], callable_coeffs='every_step')
    assert float(sol_e.evaluate()[0]) == pytest.approx(l1, rel=1e-6)
    state['k'] = 4.0
    assert float(sol_e.evaluate()[0]) == pytest.approx(l4, rel=1e-6)
    w = torch.ones(1, device='cuda:0', requires_grad=True)
    with pytest.raises(UnsupportedProblem, match='closes over trainable state'):
        build(lambda g: w * g[:, 0])


def test_tensor_core_path_refuses_unsupported_net(cuda_default):
    g = load_golden('burgers_autograd_4h', 'float64')
    with pytest.raises(RuntimeError, match='tcgen05 path needs'):
        fused('burgers_autograd_4h', g['weights'], impl=2)


MAT_CASES = sorted(k for k in problems.ZOO if 'mat' in k)


def fused_mat(name, weights):
    prob = problems.ZOO[name](tdb, 'float32')
    u = torch.as_tensor(weights, dtype=torch.float32).reshape(prob.mat_shape).to('cuda:0')
    model = tdb.Model(u, prob.domain, prob.equation, prob.conditions)
    model.compile('mat', **prob.compile_kwargs)
    return prob, model.solution_cls


@pytest.mark.parametrize('name', MAT_CASES)
def test_mat_loss_and_gradient_match_reference(name, cuda_default):
    g = load_golden(name, 'float64')
    prob, sol = fused_mat(name, g['weights'])
    sol.model.requires_grad_()
    loss, loss_n = sol.evaluate()
    loss.backward()
    grad = sol.model.grad.reshape(-1).double().cpu().numpy()
    assert float(loss) == pytest.approx(float(g['loss']), rel=LOSS_RTOL)
    assert float(loss_n) == pytest.approx(float(g['loss_normalized']), rel=LOSS_RTOL)
    gn = np.linalg.norm(g['grad'])
    assert abs(np.linalg.norm(grad) - gn) <= GRADNORM_RTOL * gn
    assert np.linalg.norm(grad - g['grad']) <= GRADVEC_RTOL * gn
    if not prob.compile_kwargs.get('tol', 0):      # (causal loss: the operator slot holds the WEIGHTED mean square)
        np.testing.assert_allclose(sol.op_mse.cpu().numpy(), g['op_mse'], rtol=1e-4)
    np.testing.assert_allclose(sol.bval_mse.cpu().numpy(), g['bval_mse'], rtol=1e-4)
    assert sol.bval_keys == [str(k) for k in g['bval_keys']]
    assert sol.bval_length == [int(x) for x in g['bval_length']]
    op = sol.op.cpu().double().numpy()
    scale = np.abs(g['op_head']).max()
    np.testing.assert_allclose(op[:256], g['op_head'], atol=2e-4 * scale, rtol=1e-3)
    bscale = np.abs(g['bval']).max()
    np.testing.assert_allclose(sol.bval.cpu().numpy(), g['bval'], atol=2e-4 * bscale, rtol=1e-4)
    np.testing.assert_allclose(sol.true_bval.cpu().numpy(), g['true_bval'], atol=1e-6)


@pytest.mark.parametrize('n,amp,gtol', [(127, 0.05, 2e-3), (511, 0.05, 2e-3), (4095, 2.0, 5e-2)])
def test_mat_large_grid_properties(n, amp, gtol, cuda_default):
    """128 x 128, 512 x 512 (interior tiles) and 4096 x 4096 Poisson (BASELINE config 4, full size).

    (i) the kernel against an independent fp64 evaluation of the same discrete operator on the device (dense banded
    D^2 matrices built by mat.derivative_band, which the CPU tests pin to the oracle);
    (ii) size-independent properties: the loss is quadratic and the gradient affine along a direction.
    At h = 1/4095 fp32 second differences carry O(1) rounding noise per cell (u * eps / 4h^2) - in the reference
    too - so (ii) uses steps large enough for the signal to dominate and (i) allows for the noise floor."""
    from torch_de_solver_b200.mat import derivative_band
    from test_mat_cpu import dense_from_band
    prob = problems.poisson_mat(tdb, 'float32', n=n)
    x = torch.linspace(0, 1, n + 1)
    u = (torch.sin(np.pi * x)[:, None] * torch.sin(np.pi * x)[None, :]).reshape(1, n + 1, n + 1).contiguous()
    model = tdb.Model(u, prob.domain, prob.equation, prob.conditions)
    model.compile('mat', **prob.compile_kwargs)
    plan = model.solution_cls._plan
    v = (torch.sin(2 * np.pi * x)[:, None] * torch.cos(3 * np.pi * x)[None, :]).reshape(1, n + 1, n + 1).contiguous()
    u0 = (u + amp * torch.sin(5 * np.pi * x)[:, None] * torch.sin(np.pi * x)[None, :]).contiguous()

    # (i) fp64 restatement: r = D2 u + u D2^T - f,  loss = lam_op mean(r^2) + lam_b/len sum_bc (u - t)^2
    h = float(x[1] - x[0])
    band, b, E = derivative_band(n + 1, 2, 2, h)
    D2 = torch.as_tensor(dense_from_band(band.astype(np.float64), b, E, n + 1), dtype=torch.float64, device='cuda:0')
    f64 = (-2 * np.pi ** 2 * torch.sin(np.pi * x.double())[:, None] * torch.sin(np.pi * x.double())[None, :])
    U = u0[0].double()
    r = D2 @ U + U @ D2.T - f64
    N = float((n + 1) ** 2)
    tgt = torch.zeros_like(U)
    tgt[:, -1] = torch.sin(np.pi * x.double())                       # u(x, y = 1) = sin(pi x); other edges 0
    mask = torch.zeros_like(U)
    cnt = torch.zeros_like(U)                                        # corners appear in two conditions
    for sl in ((0, slice(None)), (-1, slice(None)), (slice(None), 0), (slice(None), -1)):
        cnt[sl] += 1
    bdiff = (U - tgt)
    # the two conditions that share a corner have different targets only at (x=0|1, y=1): sin(pi x) = 0 there
    n_b = 4 * (n + 1)
    loss_ref = float((r ** 2).sum() / N + 100.0 * (cnt * bdiff ** 2).sum() / n_b)
    grad_ref = 2.0 / N * (D2.T @ r + r @ D2) + 100.0 * 2.0 / n_b * cnt * bdiff
    out, grad = plan.loss_grad_raw(u0)
    noise = 1.0                                                      # O(1) per-cell residual noise, see docstring
    assert float(out[0]) == pytest.approx(loss_ref, rel=1e-3, abs=(2 * noise if n > 1000 else 1e-3))
    gerr = float((grad[0].double() - grad_ref).norm() / grad_ref.norm())
    D32, U32 = D2.float(), u0[0]
    r32 = D32 @ U32 + U32 @ D32.T - f64.float()
    g32 = 2.0 / N * (D32.T @ r32 + r32 @ D32) + (100.0 * 2.0 / n_b * cnt * bdiff).float()
    floor = float((g32.double() - grad_ref).norm() / grad_ref.norm())        # fp32 rounding floor of this problem
    assert gerr < max(gtol, 3.0 * floor), (gerr, floor)

    # (ii) quadratic / affine structure with large steps (well conditioned: every difference is between values of
    # comparable magnitude; the derivative is checked at the midpoint, where the central difference of a quadratic
    # is exact)
    ts = (-50.0, 0.0, 50.0, 100.0)
    l = [float(plan.loss_grad_raw((u0 + t * v).contiguous())[0][0]) for t in ts]
    third = l[3] - 3 * l[2] + 3 * l[1] - l[0]
    assert abs(third) <= 2e-3 * max(abs(q) for q in l)
    grads = [plan.loss_grad_raw((u0 + t * v).contiguous())[1].double() for t in (0.0, 50.0, 100.0)]
    d1, d2 = grads[1] - grads[0], grads[2] - grads[1]
    if n < 1000:        # at h = 1/4095 the fp32 rounding noise of the operator itself exceeds this signal, see (i)
        assert float((d1 - d2).norm()) <= 1e-2 * float(d1.norm())
    fd = (l[3] - l[1]) / 100.0
    assert fd == pytest.approx(float((grads[1] * v.double()).sum()), rel=2e-2)


@pytest.mark.parametrize('case', ['poisson_p2_64x64', 'poisson_p2_40x132', 'poisson_p2_200x260', 'poisson_p3_72x136',
                                  'heat_p2_96x128', 'heat_p2_33x260'])
def test_mat_specialised_kernels_agree(case, cuda_default, monkeypatch):
    """The register-marching kernel, the persistent TMA kernel, the vectorised cross-stencil kernel, the register-tap kernel and the generic tiled kernel evaluate the same
    loss and gradient (interior tiles, all four kinds of boundary tiles, partial tiles)."""
    kind, p, shape = case.split('_')
    n0, n1 = (int(x) for x in shape.split('x'))
    dp = int(p[1])
    if kind == 'poisson':
        prob = problems.poisson_mat(tdb, 'float32', n=n0 - 1, ny=n1 - 1, derivative_points=dp)
    else:
        prob = problems.heat_mat(tdb, 'float32', n=n0 - 1, nt=n1 - 1, derivative_points=dp)
    u = torch.as_tensor(np.random.default_rng(3).random(prob.mat_shape, dtype=np.float32)).to('cuda:0').contiguous()
    res = {}
    all_env = ('TDB200_MAT_NO_MARCH', 'TDB200_MAT_NO_TMA', 'TDB200_MAT_NO_CROSS', 'TDB200_MAT_NO_LIN1')
    for tag, envs in (('cross-march', ()), ('cross-tma', all_env[:1]), ('cross-vec4', all_env[:2]),
                      ('register-tap', all_env[2:3]), ('generic', all_env[3:])):
        for e in all_env:
            monkeypatch.delenv(e, raising=False)
        for e in envs:
            monkeypatch.setenv(e, '1')
        model = tdb.Model(u.clone(), prob.domain, prob.equation, prob.conditions)
        model.compile('mat', **prob.compile_kwargs)
        plan = model.solution_cls._plan
        expect = tag
        if dp == 3 and tag == 'register-tap':
            expect = 'generic'                               # p = 3: > 16 taps
        if dp == 3 and tag == 'cross-march':
            expect = 'cross-tma'                             # p = 3: reach 4 keeps too many rows in registers
        assert plan.kernel_kind == expect
        out, grad = plan.loss_grad_raw(u)
        out2, grad2 = plan.loss_grad_raw(u)                  # the fused finalize step re-arms itself
        assert torch.equal(grad, grad2) and float(out[0]) == pytest.approx(float(out2[0]), rel=1e-6)
        res[tag] = (out.double().cpu().numpy(), grad.double().cpu().numpy())
    ref_out, ref_grad = res['generic']
    for tag in ('cross-march', 'cross-tma', 'cross-vec4', 'register-tap'):
        out, grad = res[tag]
        np.testing.assert_allclose(out, ref_out, rtol=2e-5)
        assert np.abs(grad - ref_grad).max() <= 2e-4 * np.abs(ref_grad).max(), tag


@pytest.mark.parametrize('case,world', [('poisson_p2_200x260', 2), ('poisson_p2_200x260', 3), ('heat_p2_96x128', 2),
                                        ('poisson_p3_72x136', 2)])
def test_mat_slab_plans_add_up(case, world, cuda_default):
    """Slab decomposition on one GPU: the plans of ranks 0..world-1 (extended slabs, loss window, boundary rows of the
    owned block) evaluated one after the other give partial loss terms that add up to, and gradient slabs that
    concatenate to, the single-plan result.  (The halo exchange itself is covered by the gloo test on the CPU.)"""
    kind, p, shape = case.split('_')
    n0, n1 = (int(x) for x in shape.split('x'))
    dp = int(p[1])
    if kind == 'poisson':
        prob = problems.poisson_mat(tdb, 'float32', n=n0 - 1, ny=n1 - 1, derivative_points=dp)
    else:
        prob = problems.heat_mat(tdb, 'float32', n=n0 - 1, nt=n1 - 1, derivative_points=dp)
    u = torch.as_tensor(np.random.default_rng(5).random(prob.mat_shape, dtype=np.float32)).to('cuda:0').contiguous()
    model = tdb.Model(u.clone(), prob.domain, prob.equation, prob.conditions)
    model.compile('mat', **prob.compile_kwargs)
    ref_out, ref_grad = model.solution_cls._plan.loss_grad_raw(u)
    outs, grads = [], []
    for rank in range(world):
        from torch_de_solver_b200.mat import slab_rows
        r0, r1 = slab_rows(n0, rank, world)
        m = tdb.Model(u[:, r0:r1].contiguous(), prob.domain, prob.equation, prob.conditions)
        m.compile('mat', **prob.compile_kwargs, shard=(rank, world))
        plan = m.solution_cls._plan
        e0, e1 = plan.ir.ext
        out, grad = plan.loss_grad_ext(u[:, e0:e1].contiguous())
        assert tuple(grad.shape) == (1, r1 - r0, n1)
        outs.append(out.double())
        grads.append(grad)
    out = torch.stack(outs).sum(0)
    np.testing.assert_allclose(out.cpu().numpy(), ref_out.double().cpu().numpy(), rtol=2e-5)
    grad = torch.cat(grads, 1)
    assert float((grad - ref_grad).abs().max()) <= 2e-4 * float(ref_grad.abs().max())


def test_repeatable_and_param_update(cuda_default):
    """Two calls give bit-identical results (fixed reduction order); weights are re-read every call."""
    g = load_golden('burgers_NN_small', 'float64')
    prob, net, sol = fused('burgers_NN_small', g['weights'])
    a = sol._plan.loss_grad().clone()
    b = sol._plan.loss_grad().clone()
    assert torch.equal(a, b)
    with torch.no_grad():
        for p in net.parameters():
            p.mul_(1.01)
    c = sol._plan.loss_grad()
    assert not torch.equal(a, c)


def test_gradient_is_derivative_of_loss(cuda_default):
    """Size-independent property: directional finite difference of the fused loss matches <grad, v>."""
    g = load_golden('kdv_autograd', 'float64')
    prob, net, sol = fused('kdv_autograd', g['weights'])
    params = list(net.parameters())
    out = sol._plan.loss_grad()
    grad = out[2 + sol._n_slots:].double()
    torch.manual_seed(1)
    v = [torch.randn_like(p) for p in params]
    vflat = torch.cat([x.reshape(-1) for x in v]).double()
    eps = 1e-3
    losses = []
    for s in (+1, -1):
        with torch.no_grad():
            for p, d in zip(params, v):
                p.add_(s * eps * d)
        losses.append(float(sol._plan.loss_grad()[0]))
        with torch.no_grad():
            for p, d in zip(params, v):
                p.sub_(s * eps * d)
    fd = (losses[0] - losses[1]) / (2 * eps)
    assert fd == pytest.approx(float(grad @ vflat), rel=2e-3)


def test_large_grid_consistency(cuda_default):
    """Full-size property (10^6 points): the loss of a grid equals the row-weighted mean of the losses of its two
    halves, and the training loop runs."""
    prob = problems.wave(tdb, 'float32', n=999, mode='autograd')
    net = problems.make_net(prob.net_layers, torch.float32, prob.init).to('cuda:0')
    model = tdb.Model(net, prob.domain, prob.equation, prob.conditions)
    model.compile(prob.mode, **prob.compile_kwargs)
    sol = model.solution_cls
    full = sol._plan.loss_grad().double()
    parts = []
    for r in range(2):
        m = tdb.Model(net, prob.domain, prob.equation, prob.conditions)
        m.compile(prob.mode, **prob.compile_kwargs, shard=(r, 2))
        parts.append(m.solution_cls._plan.loss_grad().double())
    tot = parts[0] + parts[1]
    assert float(tot[0]) == pytest.approx(float(full[0]), rel=1e-5)
    k = 2 + sol._n_slots
    gn = float(full[k:].norm())
    assert float((tot[k:] - full[k:]).norm()) <= 1e-4 * gn


def test_training_reduces_loss(cuda_default):
    g = load_golden('burgers_NN_small', 'float64')
    prob = problems.ZOO['burgers_NN_small'](tdb, 'float32')
    net = problems.make_net(prob.net_layers, torch.float32, prob.init).to('cuda:0')
    model = tdb.Model(net, prob.domain, prob.equation, prob.conditions)
    model.compile(prob.mode, **prob.compile_kwargs)
    l0 = float(model.solution_cls.evaluate()[0])
    model.train(tdb.Optimizer('Adam', {'lr': 1e-3}), 60)
    l1 = float(model.solution_cls.evaluate()[0])
    assert l1 < 0.7 * l0


def test_solution_accepts_foreign_equation_object(cuda_default):
    """Solution only reads .operator / .bconds / .h from the equation object."""
    from types import SimpleNamespace
    g = load_golden('burgers_NN_small', 'float64')
    prob = problems.ZOO['burgers_NN_small'](tdb, 'float32')
    net = problems.make_net(prob.net_layers, torch.float32, prob.init)
    set_weights(list(net.parameters()), g['weights'])
    net = net.to('cuda:0')
    grid = prob.domain.build('NN')
    bconds = prob.conditions.build(prob.domain.variable_dict)
    foreign = SimpleNamespace(grid=grid, operator=prob.equation.equation_lst, bconds=bconds, h=0.001,
                              inner_order='1', boundary_order='2')
    sol = tdb.Solution(grid, foreign, net, 'NN', None, 1, 10)
    loss, _ = sol.evaluate()
    assert float(loss) == pytest.approx(float(g['loss']), rel=LOSS_RTOL)


def test_c_abi_errors(cuda_default):
    from torch_de_solver_b200 import _native
    lib = _native.load()
    net = _native.NetDesc()
    net.n_layers = 1
    h = C.c_void_p()
    rc = lib.tdb200_plan_create(C.byref(net), 0, None, 0, None, 0, None, 0, None, 1, 0, C.byref(h))
    assert rc < 0 and lib.tdb200_last_error()


def test_cpu_device_raises():
    prob = problems.ZOO['burgers_NN_small'](tdb, 'float32')
    net = problems.make_net(prob.net_layers)
    model = tdb.Model(net, prob.domain, prob.equation, prob.conditions)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        model.compile(prob.mode, **prob.compile_kwargs)


# ---- weak form + vector-Jacobian mode ---------------------------------------------------------------------
@pytest.mark.parametrize('name', WEAK_CASES)
@pytest.mark.parametrize('impl', [0, 2])
def test_weak_form_loss_and_gradient_match_reference(name, impl, cuda_default):
    """Weak-form loss (losses.py:184-228): fields from a forward launch, nested integrals on the device, parameter
    gradient by the fused kernel in seed mode (tdb200_plan_set_field_seeds); impl=2 runs the interior on tcgen05."""
    g = load_golden(name, 'float64')
    prob = problems.ZOO[name](tdb, 'float32')
    prob, net, sol = fused(name, g['weights'], impl=impl)
    loss, loss_n = sol.evaluate()
    assert tuple(loss.shape) == (1, 1)
    loss.backward()
    grad = torch.cat([p.grad.reshape(-1) for p in net.parameters()]).double().cpu().numpy()
    nn_mode = prob.mode == 'NN'
    assert float(loss) == pytest.approx(float(g['loss']), rel=2e-4 if nn_mode else LOSS_RTOL)
    assert float(loss_n) == pytest.approx(float(g['loss_normalized']), rel=2e-4 if nn_mode else 2e-5)
    gn = np.linalg.norm(g['grad'])
    assert abs(np.linalg.norm(grad) - gn) <= (5e-4 if nn_mode else GRADNORM_RTOL) * gn
    assert np.linalg.norm(grad - g['grad']) <= (5e-4 if nn_mode else GRADVEC_RTOL) * gn


@pytest.mark.parametrize('name', ['kdv_autograd', 'navier_stokes_autograd', 'burgers_NN_small', 'legendre_autograd'])
def test_vector_jacobian_product_of_fields(name, cuda_default):
    """Operator.operator_compute() / Bounds.apply_bcs() are differentiable: an arbitrary function of the per-point
    fields back-propagates through one fused launch in seed mode; checked against torch autograd through the oracle."""
    g = load_golden(name, 'float64')
    prob, net, sol = fused(name, g['weights'])
    op = sol.operator.operator_compute()
    bval, tval, keys, lens = sol.boundary.apply_bcs()
    assert op.requires_grad and bval.requires_grad
    gen = torch.Generator(device='cpu').manual_seed(5)
    c_op = torch.randn(op.shape, generator=gen, dtype=torch.float64, device='cpu')
    c_b = torch.randn(bval.shape, generator=gen, dtype=torch.float64, device='cpu')
    f = (op * torch.tanh(op) * c_op.to(op)).sum() + (torch.sin(bval) * c_b.to(bval)).sum()
    f.backward()
    grad = torch.cat([p.grad.reshape(-1) for p in net.parameters()]).double().cpu().numpy()
    # the oracle, fp64 on the CPU
    torch.set_default_device('cpu')
    try:
        osol, _, _, _ = oracle_eval(name, 'float64', g['weights'])
        osol.evaluate()                                   # fresh graph for op / bval
        f_o = (osol.op * torch.tanh(osol.op) * c_op).sum() + (torch.sin(osol.bval) * c_b).sum()
        grads_o = torch.autograd.grad(f_o, list(osol.model.parameters()))
    finally:
        torch.set_default_device('cuda:0')
    grad_o = torch.cat([x.reshape(-1) for x in grads_o]).numpy()
    nn_mode = prob.mode == 'NN'
    assert float(f) == pytest.approx(float(f_o), rel=2e-3 if nn_mode else 2e-5)
    assert np.linalg.norm(grad - grad_o) <= (5e-3 if nn_mode else 2e-4) * np.linalg.norm(grad_o)
    # and the plain loss path is untouched by the seed-mode call
    loss, _ = sol.evaluate()
    assert float(loss) == pytest.approx(float(g['loss']), rel=LOSS_RTOL)


def test_mat_stencil_timing_entry_points(cuda_default):
    """Measurement aids of the C ABI (bench.py roofline): events around the stencil kernel of an eager call, and
    back-to-back launches of the stencil kernel alone; neither changes the results of later calls."""
    prob = problems.poisson_mat(tdb, 'float32', n=255, ny=255, derivative_points=2)
    u = problems.make_mat_model(prob.mat_shape, torch.float32).to('cuda:0').contiguous()
    model = tdb.Model(u, prob.domain, prob.equation, prob.conditions)
    model.compile('mat', **prob.compile_kwargs)
    plan = model.solution_cls._plan
    assert plan.kernel_kind == 'cross-march' and plan.launches_per_call in (1, 2)    # 1: the single-launch step
    out0, grad0 = plan.loss_grad_raw(u)
    plan.set_timing(True)
    plan.loss_grad_raw(u)
    assert 0.0 < plan.stencil_ms() < 50.0
    plan.set_timing(False)
    assert 0.0 < plan.time_stencil(u, 5) < 50.0
    out1, grad1 = plan.loss_grad_raw(u)
    assert torch.equal(grad0, grad1) and float(out0[0]) == pytest.approx(float(out1[0]), rel=1e-6)


def test_lambda_changes_are_picked_up(cuda_default):
    """Callbacks replace or modify lambda_operator / lambda_bound between steps (AdaptiveLambda): new objects, new
    values of the same kind and in-place updates must all reach the plan (the steady state re-reads nothing)."""
    g = load_golden('burgers_NN_small', 'float64')
    prob, net, sol = fused('burgers_NN_small', g['weights'])

    def check(lam_op, lam_b):
        loss, _ = sol.evaluate()
        want = float(sol.op_mse.sum()) * lam_op + float(sol.bval_mse.sum()) * lam_b
        assert float(loss) == pytest.approx(want, rel=1e-5)

    check(1.0, 10.0)
    sol.lambda_bound = 3.0
    check(1.0, 3.0)
    for v in (7.0, 9.0, 11.0):                       # fresh tensors: a recycled id must not look unchanged
        sol.lambda_bound = torch.tensor([[v]], device='cuda:0')
        check(1.0, v)
    sol.lambda_bound.mul_(2.0)                       # in place
    check(1.0, 22.0)
    sol.lambda_operator = torch.tensor([[0.5]], device='cuda:0')
    check(0.5, 22.0)


def test_graph_captured_step_matches_eager(cuda_default):
    """FusedPlan.capture(): the replayed graph (fork / join of the boundary launches included) returns the eager result
    and follows in-place parameter updates."""
    g = load_golden('burgers_NN_cfg1', 'float64')
    prob, net, sol = fused('burgers_NN_cfg1', g['weights'])
    eager = sol._plan.loss_grad().clone()
    replay, out = sol._plan.capture()
    replay()
    assert torch.equal(out, eager)
    with torch.no_grad():
        for p in net.parameters():
            p.mul_(1.01)
    replay()
    assert torch.equal(out, sol._plan.loss_grad())
    assert not torch.equal(out, eager)


# ---- fused optimiser + graph-captured training step ------------------------------------------------------
@pytest.mark.parametrize('opt_name,opt_kw', [('Adam', dict(lr=1e-3)), ('AdamW', dict(lr=2e-3, weight_decay=0.05)),
                                             ('SGD', dict(lr=1e-4, momentum=0.9, weight_decay=1e-4)),
                                             ('Adam', dict(lr=1e-3, weight_decay=1e-3, betas=(0.8, 0.99)))])
def test_fused_optimizer_matches_torch(opt_name, opt_kw, cuda_default):
    """50 steps of tdb200_optimizer_step against torch.optim on the same gradients (the fused plan's own)."""
    from torch_de_solver_b200.optimizers.fused import FusedOptimizer
    g = load_golden('burgers_NN_small', 'float64')
    prob, net_a, sol_a = fused('burgers_NN_small', g['weights'])
    prob, net_b, sol_b = fused('burgers_NN_small', g['weights'])
    topt = getattr(torch.optim, opt_name)(net_a.parameters(), **opt_kw)
    fopt = FusedOptimizer(opt_name, sol_b._ir.net.param_tensors(), **opt_kw)
    for _ in range(50):
        topt.zero_grad()
        loss, _ = sol_a.evaluate()
        loss.backward()
        topt.step()
        out, flat = sol_b._run_plan()
        fopt.step(flat)
    wa = torch.cat([p.detach().reshape(-1) for p in net_a.parameters()])
    wb = torch.cat([p.detach().reshape(-1) for p in net_b.parameters()])
    assert int(fopt.step_count) == 50
    assert float((wa - wb).norm()) <= 2e-6 * float(wa.norm())
    assert float(sol_b.evaluate()[0]) == pytest.approx(float(sol_a.evaluate()[0]), rel=1e-4)


@pytest.mark.parametrize('mode_name', ['burgers_NN_small', 'wave_autograd', 'poisson_mat_p2_rect'])
def test_model_train_uses_graph_step_and_matches_eager(mode_name, cuda_default, monkeypatch):
    """Model.train (tedeous/model.py:134-195) through the captured training step = the eager loop with torch.optim."""
    g = load_golden(mode_name, 'float64')

    def run(eager):
        if eager:
            monkeypatch.setenv('TDB200_EAGER_TRAIN', '1')
        else:
            monkeypatch.delenv('TDB200_EAGER_TRAIN', raising=False)
        prob = problems.ZOO[mode_name](tdb, 'float32')
        if prob.mode == 'mat':
            net = torch.as_tensor(g['weights']).reshape(prob.mat_shape).float().to('cuda:0').contiguous()
        else:
            net = problems.make_net(prob.net_layers, torch.float32, prob.init)
            set_weights(list(net.parameters()), g['weights'])
            net = net.to('cuda:0')
        model = tdb.Model(net, prob.domain, prob.equation, prob.conditions)
        model.compile(prob.mode, **prob.compile_kwargs)
        seen = []

        class Probe:
            def set_model(self, m): self.model = m
            def on_epoch_end(self, logs=None): seen.append(float(self.model.cur_loss))
        model.train(tdb.Optimizer('Adam', {'lr': 1e-3}, gamma=0.9, decay_every=10), 31, callbacks=[Probe()])
        w = model.solution_cls.model if prob.mode == 'mat' else torch.cat([p.detach().reshape(-1) for p in model.net.parameters()])
        return seen, w.detach().reshape(-1).clone(), model
    seen_e, w_e, _ = run(True)
    seen_f, w_f, model = run(False)
    assert model._fused_train_step(tdb.Optimizer('Adam', {'lr': 1e-3}), False) is not None
    assert len(seen_e) == len(seen_f) == 30
    np.testing.assert_allclose(seen_f, seen_e, rtol=2e-4)
    assert seen_f[-1] < seen_f[0]
    assert float((w_e - w_f).norm()) <= 2e-5 * float(w_e.norm())


def test_mini_batches_match_reference(cuda_default):
    """batch_size in mode 'autograd' (tedeous/eval.py:124-141, 174-182): seven consecutive steps - five batches of the
    first epoch (the last one ragged: 41 rows), a reshuffle, two of the next - against the fixture of the unmodified
    reference (tests/golden/make_minibatch.py), with the same generator state as the reference's DataLoader."""
    import os
    from helpers import GOLDEN_DIR
    g = dict(np.load(os.path.join(GOLDEN_DIR, 'minibatch_wave.npz')))
    prob = problems.wave(tdb, 'float32', n=20, mode='autograd', layers=(2, 32, 32, 1))
    net = problems.make_net(prob.net_layers, torch.float32, prob.init)
    set_weights(list(net.parameters()), g['weights'])
    net = net.to('cuda:0')
    model = tdb.Model(net, prob.domain, prob.equation, prob.conditions, batch_size=100)
    model.compile('autograd', **prob.compile_kwargs, batch_generator=torch.Generator())     # CPU generator, default seed
    sol = model.solution_cls
    assert sol.operator.n_batches == int(g['n_batches']) == 5
    for i in range(7):
        for p in net.parameters():
            p.grad = None
        loss, _ = sol.evaluate()
        loss.backward()
        grad = torch.cat([p.grad.reshape(-1) for p in net.parameters()]).double().cpu().numpy()
        assert float(loss) == pytest.approx(float(g['losses'][i]), rel=LOSS_RTOL), i
        gn = np.linalg.norm(g['grads'][i])
        assert np.linalg.norm(grad - g['grads'][i]) <= GRADVEC_RTOL * gn, i
        assert sol.operator.current_batch_i == (i + 1) % 5
    assert sol.save_op.shape[0] == 200                  # two batches into the second epoch
    # NN mode: the reference's batches never reach the operator (q8) -> batch_size changes nothing
    g2 = load_golden('burgers_NN_small', 'float64')
    prob2 = problems.ZOO['burgers_NN_small'](tdb, 'float32')
    net2 = problems.make_net(prob2.net_layers, torch.float32, prob2.init)
    set_weights(list(net2.parameters()), g2['weights'])
    m2 = tdb.Model(net2.to('cuda:0'), prob2.domain, prob2.equation, prob2.conditions, batch_size=64)
    m2.compile('NN', **prob2.compile_kwargs)
    assert float(m2.solution_cls.evaluate()[0]) == pytest.approx(float(g2['loss']), rel=LOSS_RTOL)


@pytest.mark.parametrize('name', ['kdv_autograd', 'navier_stokes_autograd', 'burgers_NN_small', 'wave_autograd',
                                  'burgers_inverse_autograd'])
def test_residual_jacobian_rows(name, cuda_default):
    """per-residual Jacobian rows (tdb200_jacobian_rows) against torch autograd through the oracle - sampled
    rows one by one (what NGD.gram_factory does, tedeous/optimizers/ngd.py:57-77), J^T c against one reverse sweep, and
    J v against the directional derivative of the oracle's residuals."""
    g = load_golden(name, 'float64')
    prob, net, sol = fused(name, g['weights'])
    j_op, j_bnd = sol.residual_jacobian()
    op, bval = sol.op, sol.bval
    assert j_op.shape == (op.numel(), sol._plan.n_params) and j_bnd.shape == (bval.numel(), sol._plan.n_params)
    j_op, j_bnd = j_op.double().cpu(), j_bnd.double().cpu()
    gen = torch.Generator(device='cpu').manual_seed(11)
    rows_op = torch.randint(0, j_op.shape[0], (6,), generator=gen, device='cpu').tolist()
    rows_b = torch.randint(0, j_bnd.shape[0], (6,), generator=gen, device='cpu').tolist()
    c_op = torch.randn(j_op.shape[0], generator=gen, dtype=torch.float64, device='cpu')
    c_b = torch.randn(j_bnd.shape[0], generator=gen, dtype=torch.float64, device='cpu')
    v = torch.randn(j_op.shape[1], generator=gen, dtype=torch.float64, device='cpu')
    jv_op, jv_b = sol.residual_jvp(v)
    torch.set_default_device('cpu')
    try:
        osol, _, _, _ = oracle_eval(name, 'float64', g['weights'])
        osol.evaluate()
        params = list(osol.model.parameters())
        r_op, r_b = osol.op.reshape(-1), (osol.bval - osol.true_bval).reshape(-1)

        def flat_grad(scalar):
            gs = torch.autograd.grad(scalar, params, retain_graph=True, allow_unused=True)
            return torch.cat([torch.zeros_like(p).reshape(-1) if x is None else x.reshape(-1) for x, p in zip(gs, params)])
        nn_mode = prob.mode == 'NN'
        tol = 5e-3 if nn_mode else 2e-4
        for r in rows_op:
            ref = flat_grad(r_op[r])
            assert torch.linalg.norm(j_op[r] - ref) <= tol * max(float(torch.linalg.norm(ref)), 1e-12)
        for r in rows_b:
            ref = flat_grad(r_b[r])
            assert torch.linalg.norm(j_bnd[r] - ref) <= tol * float(torch.linalg.norm(ref)) + 1e-12
        ref = flat_grad((c_op * r_op).sum() + (c_b * r_b).sum())
        got = j_op.T @ c_op + j_bnd.T @ c_b
        assert torch.linalg.norm(got - ref) <= tol * float(torch.linalg.norm(ref))
        # J v: central difference of the oracle's residuals along v (fp64, step 1e-6)
        base = [p.detach().clone() for p in params]

        def residuals_at(eps):
            off = 0
            with torch.no_grad():
                for p, b in zip(params, base):
                    p.copy_(b + eps * v[off:off + p.numel()].reshape(p.shape))
                    off += p.numel()
            osol.evaluate()
            return osol.op.reshape(-1).detach().clone(), (osol.bval - osol.true_bval).reshape(-1).detach().clone()
        (a0, b0), (a1, b1) = residuals_at(-1e-6), residuals_at(1e-6)
        fd_op, fd_b = (a1 - a0) / 2e-6, (b1 - b0) / 2e-6
    finally:
        torch.set_default_device('cuda:0')
    assert torch.linalg.norm(jv_op.double().cpu() - fd_op) <= max(tol, 1e-3) * float(torch.linalg.norm(fd_op))
    assert torch.linalg.norm(jv_b.double().cpu() - fd_b) <= max(tol, 1e-3) * float(torch.linalg.norm(fd_b)) + 1e-12
    # the plain loss path is untouched
    loss, _ = sol.evaluate()
    assert float(loss) == pytest.approx(float(g['loss']), rel=LOSS_RTOL)


def test_ngd_step_reduces_loss(cuda_default):
    """Model.train with the natural-gradient optimiser (tedeous/optimizers/ngd.py) on the per-residual Jacobian rows of
    the fused path: G = J^T J / n is symmetric and its pseudo-inverse solve reproduces right-hand sides in its range and three epochs reduce the loss (grid line search: never uphill)."""
    prob = problems.ZOO['wave_autograd'](tdb, 'float32')
    torch.manual_seed(3)
    net = torch.nn.Sequential(torch.nn.Linear(2, 16), torch.nn.Tanh(), torch.nn.Linear(16, 16), torch.nn.Tanh(),
                              torch.nn.Linear(16, 1)).to('cuda:0')
    model = tdb.Model(net, prob.domain, prob.equation, prob.conditions)
    model.compile(prob.mode, **prob.compile_kwargs)
    sol = model.solution_cls
    loss0, _ = sol.evaluate()
    from torch_de_solver_b200.optimizers.ngd import NGD
    j_op, j_bnd = sol.residual_jacobian()
    G = NGD.gram(j_op) + NGD.gram(j_bnd)
    assert torch.allclose(G, G.T, atol=1e-6 * float(G.abs().max()))
    b = G @ torch.randn(G.shape[0], device=G.device)                  # a right-hand side in the range of G
    x = NGD.pinv_solve(G.double(), b.double(), tol=1e-10 * float(torch.linalg.matrix_norm(G.double(), 2)))
    assert torch.linalg.norm(G.double() @ x - b.double()) <= 1e-4 * torch.linalg.norm(b.double())
    model.train(tdb.Optimizer('NGD', {'grid_steps_number': 20}), epochs=4, info_string_every=None)
    loss1, _ = sol.evaluate()
    print('NGD: loss', float(loss0), '->', float(loss1))
    assert torch.isfinite(loss1).all() and float(loss1) < float(loss0)
