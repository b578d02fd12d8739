"""Problem zoo shared by the golden generator (reference API), the oracle tests and the GPU parity tests.

Every builder takes an `api` namespace exposing Domain / Conditions / Equation (either the reference's
`tedeous.data` or `torch_de_solver_b200`) and a dtype string, and returns a `Problem`.  The operator dicts
follow the reference's example scripts (file:line in each docstring)."""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, List

import numpy as np
import torch


@dataclass
class Problem:
    name: str
    domain: object
    conditions: object
    equation: object
    mode: str
    net_layers: List[int]            # [d, w, ..., n_out]  (mat mode: [])
    compile_kwargs: Dict = field(default_factory=dict)
    init: str = 'default'            # 'default' (Kaiming-uniform) | 'xavier'
    mat_shape: tuple = ()
    train_steps: int = 0             # fixture state: weights after this many reference Adam steps (lr 1e-3)


class InitWithParams(str):
    """An `init` name that also carries trainable scalar coefficients (inverse problems, tedeous/models.py:183-195
    `parameter_registr`): `make_net` registers them on the net and calls `bind(net)`, which puts the live Parameter
    objects into the equation's terms (the reference example does the same by hand:
    examples/examples_burgers/example_burgers_1d_inverse.py:59-88)."""
    def __new__(cls, name, params, bind):
        obj = super().__new__(cls, name)
        obj.params, obj.bind = dict(params), bind
        return obj


def make_net(layers: List[int], dtype=torch.float32, init='default', seed=0) -> torch.nn.Sequential:
    net = _make_net(layers, dtype, init, seed)
    if isinstance(init, InitWithParams):
        for key, value in init.params.items():
            net.register_parameter(key, torch.nn.Parameter(torch.tensor([value], dtype=dtype)))
        init.bind(net)
    return net


def _make_net(layers: List[int], dtype=torch.float32, init='default', seed=0) -> torch.nn.Sequential:
    torch.manual_seed(seed)
    mods = []
    for i in range(len(layers) - 1):
        mods.append(torch.nn.Linear(layers[i], layers[i + 1]))
        if i < len(layers) - 2:
            mods.append(torch.nn.Tanh())
    net = torch.nn.Sequential(*mods)
    if init == 'xavier':
        for m in net.modules():
            if isinstance(m, torch.nn.Linear):
                torch.nn.init.xavier_normal_(m.weight)
                torch.nn.init.zeros_(m.bias)
    elif init == 'xavier_b':       # xavier weights, small random biases (keeps all gradients non-trivial)
        for m in net.modules():
            if isinstance(m, torch.nn.Linear):
                torch.nn.init.xavier_normal_(m.weight)
                torch.nn.init.uniform_(m.bias, -0.3, 0.3)
    return net.to(dtype)


def make_mat_model(shape, dtype=torch.float32, seed=0) -> torch.Tensor:
    """mat-mode "model" = the solution values on the grid ([n_eq, N0, N1], tedeous/models.py:198-226).  The
    default all-ones tensor has zero derivatives everywhere, so tests use a smooth field plus noise."""
    torch.manual_seed(seed)
    if len(shape) == 2:          # 1-D grid
        n_eq, n0 = shape
        x = torch.linspace(0, 1, n0)
        base = torch.stack([torch.sin(3 * x + k) for k in range(n_eq)])
        return (base + 0.05 * torch.rand(shape)).to(dtype)
    n_eq, n0, n1 = shape
    x = torch.linspace(0, 1, n0)[:, None]
    y = torch.linspace(0, 1, n1)[None, :]
    base = torch.stack([torch.sin(3 * x + k) * torch.cos(2 * y - k) for k in range(n_eq)])
    return (base + 0.05 * torch.rand(shape)).to(dtype)


# --- config 1: Burgers (examples/examples_burgers/example_burgers_1d.py:35-80) ----------------------------
def burgers(api, dtype='float32', n=100, mode='NN', layers=(2, 100, 100, 100, 1), h=0.001, n0=None, tol=0):
    mu = 0.01 / math.pi
    dom = api.Domain()
    dom.variable('x', [-1, 1], n0 or n, dtype=dtype)
    dom.variable('t', [0, 1], n, dtype=dtype)
    bc = api.Conditions()
    bc.dirichlet({'x': [-1, 1], 't': 0}, value=lambda g: -torch.sin(np.pi * g[:, 0]))
    bc.dirichlet({'x': -1, 't': [0, 1]}, value=0)
    bc.dirichlet({'x': 1, 't': [0, 1]}, value=0)
    eq = api.Equation()
    eq.add({
        'du/dt**1': {'coeff': 1., 'du/dt': [1], 'pow': 1, 'var': 0},
        '+u*du/dx': {'coeff': 1, 'u*du/dx': [[None], [0]], 'pow': [1, 1], 'var': [0, 0]},
        '-mu*d2u/dx2': {'coeff': -mu, 'd2u/dx2': [0, 0], 'pow': 1, 'var': 0},
    })
    kw = dict(lambda_operator=1, lambda_bound=10)
    if mode == 'NN':
        kw['h'] = h
    if tol:
        kw['tol'] = tol                 # causal loss (tedeous/losses.py:137-182)
    return Problem(f'burgers_{mode}', dom, bc, eq, mode, list(layers), kw)


# --- inverse problem: trainable coefficients (examples/examples_burgers/example_burgers_1d_inverse.py:59-88) -----
def burgers_inverse(api, dtype='float32', n=30, mode='autograd', layers=(2, 100, 100, 100, 1), h=0.01):
    """Burgers with the convection and diffusion coefficients as trainable scalars registered on the net, a `data`
    condition on scattered points next to the Dirichlet ones."""
    dom = api.Domain()
    dom.variable('x', [-1, 1], n, dtype=dtype)
    dom.variable('t', [0, 1], n, dtype=dtype)
    tdt = torch.float64 if dtype == 'float64' else torch.float32
    bc = api.Conditions()
    bc.dirichlet({'x': [-1, 1], 't': 0}, value=lambda g: -torch.sin(np.pi * g[:, 0]))
    bc.dirichlet({'x': -1, 't': [0, 1]}, value=0)
    gen = torch.Generator(device='cpu').manual_seed(7)
    pts = torch.rand(40, 2, generator=gen, dtype=torch.float64, device='cpu').to(torch.empty(0).device)
    pts[:, 0] = 2 * pts[:, 0] - 1
    vals = -torch.sin(np.pi * pts[:, 0]) * torch.exp(-pts[:, 1])
    bc.data(bnd=pts.to(tdt), operator=None, value=vals.to(tdt))
    eq_dict = {
        'du/dt**1': {'coeff': 1., 'du/dt': [1], 'pow': 1, 'var': 0},
        '+lam1*u*du/dx': {'coeff': None, 'u*du/dx': [[None], [0]], 'pow': [1, 1], 'var': [0, 0]},
        '-lam2*d2u/dx2': {'coeff': None, 'd2u/dx2': [0, 0], 'pow': 1, 'var': 0},
    }
    eq = api.Equation()
    eq.add(eq_dict)

    def bind(net):
        eq_dict['+lam1*u*du/dx']['coeff'] = net.lam1
        eq_dict['-lam2*d2u/dx2']['coeff'] = net.lam2
    kw = dict(lambda_operator=1, lambda_bound=100)
    if mode == 'NN':
        kw['h'] = h
    return Problem(f'burgers_inverse_{mode}', dom, bc, eq, mode, list(layers), kw,
                   init=InitWithParams('xavier_b', {'lam1': 2., 'lam2': -0.2}, bind))


# --- config 2: wave (examples/examples_wave/example_wave_1d_basic.py:36-97) ---------------------------------
def wave(api, dtype='float32', n=40, mode='autograd', layers=(2, 100, 100, 100, 1), h=0.01, operator_ic=True, n0=None):
    def exact(g):
        x, t = g[:, 0], g[:, 1]
        return torch.sin(np.pi * x) * torch.cos(2 * np.pi * t) + 0.5 * torch.sin(4 * np.pi * x) * torch.cos(8 * np.pi * t)
    dom = api.Domain()
    dom.variable('x', [0, 1], n0 or n, dtype=dtype)
    dom.variable('t', [0, 1], n, dtype=dtype)
    bc = api.Conditions()
    bc.dirichlet({'x': [0, 1], 't': 0}, value=exact)
    if operator_ic:
        bc.operator({'x': [0, 1], 't': 0}, operator={'du/dt': {'coeff': 1, 'term': [1], 'pow': 1, 'var': 0}}, value=0)
    bc.dirichlet({'x': 0, 't': [0, 1]}, value=exact)
    bc.dirichlet({'x': 1, 't': [0, 1]}, value=exact)
    eq = api.Equation()
    eq.add({
        'd2u/dt2**1': {'coeff': 1, 'd2u/dt2': [1, 1], 'pow': 1},
        '-C*d2u/dx2**1': {'coeff': -4, 'd2u/dx2': [0, 0], 'pow': 1},
    })
    kw = dict(lambda_operator=1, lambda_bound=100)
    if mode == 'NN':
        kw['h'] = h
    return Problem(f'wave_{mode}', dom, bc, eq, mode, list(layers), kw, init='xavier_b')


# --- config 3: KdV periodic (examples/examples_korteweg_de_vries/example_KdV_periodic.py:22-148) -----------
def kdv(api, dtype='float32', nx=30, nt=30, mode='autograd', layers=(2, 100, 100, 100, 1), h=0.01):
    def soliton(x, t):
        E = torch.exp(x)
        return 2 / (torch.cosh((x - 4 * t) / 1.0)) ** 2 + 0 * E
    dom = api.Domain()
    dom.variable('x', [-10, 10], nx, dtype=dtype)
    dom.variable('t', [0, 1], nt, dtype=dtype)
    bc = api.Conditions()
    bc.periodic([{'x': -10, 't': [0, 1]}, {'x': 10, 't': [0, 1]}])
    x = dom.variable_dict['x']
    bc.dirichlet({'x': [-10, 10], 't': 0}, value=soliton(x, torch.tensor([0.])))
    eq = api.Equation()
    eq.add({
        '1*du/dt**1': {'coeff': 1, 'du/dt': [1], 'pow': 1, 'var': 0},
        '6*u**1*du/dx**1': {'coeff': 6, 'u*du/dx': [[None], [0]], 'pow': [1, 1], 'var': [0, 0]},
        'd3u/dx3**1': {'coeff': 1, 'd3u/dx3': [0, 0, 0], 'pow': 1, 'var': 0},
    })
    kw = dict(lambda_operator=1, lambda_bound=100)
    if mode == 'NN':
        kw['h'] = h
    return Problem(f'kdv_{mode}', dom, bc, eq, mode, list(layers), kw)


# --- config 5: Navier-Stokes 2D+t (examples/examples_navier_stokes/example_navier_stokes_2d_long_time.py:36-198)
def navier_stokes(api, dtype='float32', n=8, layers=(3, 100, 100, 100, 100, 100, 100, 3), n0=None):
    ro, mu = 1., 1.
    A1, A2, A3 = 1., 1., 1.
    dom = api.Domain()
    dom.variable('x', [0, 5], n0 or n, dtype=dtype)
    dom.variable('y', [0, 1], n, dtype=dtype)
    dom.variable('t', [0, 5], n, dtype=dtype)
    bc = api.Conditions()
    for v in range(3):
        bc.dirichlet({'x': [0, 5], 'y': [0, 1], 't': 0}, value=0, var=v)
    inlet = lambda g: torch.sin(np.pi * g[:, 1]) * (A1 * torch.sin(np.pi * g[:, 2]) + A2 * torch.sin(3 * np.pi * g[:, 2])
                                                    + A3 * torch.sin(5 * np.pi * g[:, 2]))
    bc.dirichlet({'x': 0, 'y': [0, 1], 't': [0, 5]}, value=inlet, var=0)
    bc.dirichlet({'x': 5, 'y': [0, 1], 't': [0, 5]}, value=0, var=0)
    bc.dirichlet({'x': 0, 'y': [0, 1], 't': [0, 5]}, value=0, var=1)
    bc.dirichlet({'x': 5, 'y': [0, 1], 't': [0, 5]}, value=0, var=1)
    bc.dirichlet({'x': 5, 'y': [0, 1], 't': [0, 5]}, value=0, var=2)

    def forcing(g):
        return -torch.sin(np.pi * g[:, 0]) * torch.sin(np.pi * g[:, 1]) * torch.sin(np.pi * g[:, 2])
    eq = api.Equation()
    eq.add({
        'du/dx': {'coeff': 1, 'term': [0], 'pow': 1, 'var': 0},
        'dv/dy': {'coeff': 1, 'term': [1], 'pow': 1, 'var': 1},
    })
    eq.add({
        'du/dt': {'coeff': 1, 'term': [2], 'pow': 1, 'var': 0},
        'u * du/dx': {'coeff': 1, 'term': [[None], [0]], 'pow': [1, 1], 'var': [0, 0]},
        'v * du/dy': {'coeff': 1, 'term': [[None], [1]], 'pow': [1, 1], 'var': [1, 0]},
        '1/ro * dp/dx': {'coeff': 1 / ro, 'term': [0], 'pow': 1, 'var': 2},
        '-mu * d2u/dx2': {'coeff': -mu, 'term': [0, 0], 'pow': 1, 'var': 0},
        '-mu * d2u/dy2': {'coeff': -mu, 'term': [1, 1], 'pow': 1, 'var': 0},
    })
    eq.add({
        'dv/dt': {'coeff': 1, 'term': [2], 'pow': 1, 'var': 1},
        'u * dv/dx': {'coeff': 1, 'term': [[None], [0]], 'pow': [1, 1], 'var': [0, 1]},
        'v * dv/dy': {'coeff': 1, 'term': [[None], [1]], 'pow': [1, 1], 'var': [1, 1]},
        '1/ro * dp/dy': {'coeff': 1 / ro, 'term': [1], 'pow': 1, 'var': 2},
        '-mu * d2v/dx2': {'coeff': -mu, 'term': [0, 0], 'pow': 1, 'var': 1},
        '-mu * d2v/dy2': {'coeff': -mu, 'term': [1, 1], 'pow': 1, 'var': 1},
        '-f(x, y, t)': {'coeff': forcing, 'term': [None], 'pow': 0},
    })
    return Problem('navier_stokes_autograd', dom, bc, eq, 'autograd', list(layers),
                   dict(lambda_operator=1, lambda_bound=1000), init='xavier_b')


# --- extra coverage: ODE with operator BC, nonlinear powers, tensor / callable coefficients ----------------
def legendre_ode(api, dtype='float32', n=40, mode='autograd', layers=(1, 32, 32, 1), order=3, h=0.001):
    """(1 - t^2) u'' - 2 t u' + n(n+1) u = 0, u(0) = P_n(0), u'(1) = P_n'(1)
    (examples/examples_legendre/example_legendre.py pattern: callable coefficients, operator BC)."""
    from numpy.polynomial import legendre as L
    dom = api.Domain()
    dom.variable('t', [0, 1], n, dtype=dtype)
    t0, t1 = torch.tensor([[0.]]), torch.tensor([[1.]])
    c = [0] * order + [1]
    bc = api.Conditions()
    bc.dirichlet(t0, value=torch.tensor([float(L.legval(0., c))]))
    bc.operator(t1, operator={'du/dt': {'coeff': 1, 'du/dt': [0], 'pow': 1}},
                value=torch.tensor([float(L.legval(1., L.legder(c)))]))
    eq = api.Equation()
    eq.add({
        '(1-t^2)*d2u/dt2': {'coeff': lambda g: 1 - g[:, 0] ** 2, 'd2u/dt2': [0, 0], 'pow': 1},
        '-2t*du/dt': {'coeff': lambda g: -2 * g[:, 0], 'du/dt': [0], 'pow': 1},
        'n(n+1)u': {'coeff': float(order * (order + 1)), 'u': [None], 'pow': 1},
    })
    kw = dict(lambda_operator=1, lambda_bound=10)
    if mode == 'NN':
        kw['h'] = h
    if mode == 'mat':           # 1-D grid, model [1, n + 1] (examples/examples_legendre/example_ODE_Legendre_matrix.py)
        return Problem('legendre_mat', dom, bc, eq, 'mat', [], dict(lambda_operator=1, lambda_bound=10, derivative_points=2),
                       mat_shape=(1, n + 1))
    return Problem(f'legendre_{mode}', dom, bc, eq, mode, list(layers), kw)


def nonlinear_mix(api, dtype='float32', n=16, mode='autograd', layers=(2, 24, 24, 2), h=0.01):
    """Two-output system with powers, products across variables, 4th-order derivative, tensor coefficient,
    forcing term, data condition and a per-type lambda list."""
    dom = api.Domain()
    dom.variable('x', [0, 1], n, dtype=dtype)
    dom.variable('t', [0, 2], n, dtype=dtype)
    N = (n + 1) ** 2
    tdt = torch.float64 if dtype == 'float64' else torch.float32
    coeff_t = torch.linspace(0.5, 1.5, N, dtype=tdt)
    bc = api.Conditions()
    bc.dirichlet({'x': [0, 1], 't': 0}, value=lambda g: torch.cos(g[:, 0]), var=0)
    bc.dirichlet({'x': 0, 't': [0, 2]}, value=0.3, var=1)
    bc.operator({'x': 1, 't': [0, 2]}, operator={'dv/dx': {'coeff': 2., 'term': [0], 'pow': 1, 'var': 1},
                                                 'u': {'coeff': -1., 'term': [None], 'pow': 1, 'var': 0}}, value=0.1)
    data_pts = torch.tensor([[0.25, 0.5], [0.5, 1.0], [0.75, 1.5]], dtype=tdt)
    bc.data(data_pts, None, torch.tensor([0.1, 0.2, 0.3], dtype=tdt), var=1)
    eq = api.Equation()
    eq.add({
        'du/dt': {'coeff': 1, 'term': [1], 'pow': 1, 'var': 0},
        'c(x)*u^2*dv/dx': {'coeff': coeff_t, 'term': [[None], [0]], 'pow': [2, 1], 'var': [0, 1]},
        '-0.01*d4u/dx4': {'coeff': -0.01, 'term': [0, 0, 0, 0], 'pow': 1, 'var': 0},
        'f': {'coeff': lambda g: torch.sin(g[:, 0] + g[:, 1]), 'term': [None], 'pow': 0},
    })
    eq.add({
        'd2v/dt2': {'coeff': 1, 'term': [1, 1], 'pow': 1, 'var': 1},
        '(dv/dx)^2': {'coeff': 0.5, 'term': [0], 'pow': 2, 'var': 1},
        'u*v': {'coeff': -1.5, 'term': [[None], [None]], 'pow': [1, 1], 'var': [0, 1]},
        'd3u/dx3': {'coeff': 0.1, 'term': [0, 0, 0], 'pow': 1, 'var': 0},
    })
    kw = dict(lambda_operator=[1., 2.], lambda_bound=[10., 5., 3.])
    if mode == 'NN':
        kw['h'] = h
    return Problem(f'nonlinear_mix_{mode}', dom, bc, eq, mode, list(layers), kw, init='xavier_b')


# --- config 4: Poisson, mat mode -------------------------------------------------------
def poisson_mat(api, dtype='float32', n=32, derivative_points=2, ny=None):
    ny = n if ny is None else ny
    dom = api.Domain()
    dom.variable('x', [0, 1], n, dtype=dtype)
    dom.variable('y', [0, 1], ny, dtype=dtype)
    bc = api.Conditions()
    bc.dirichlet({'x': 0, 'y': [0, 1]}, value=0)
    bc.dirichlet({'x': 1, 'y': [0, 1]}, value=0)
    bc.dirichlet({'x': [0, 1], 'y': 0}, value=0)
    bc.dirichlet({'x': [0, 1], 'y': 1}, value=lambda g: torch.sin(np.pi * g[:, 0]))
    tdt = torch.float64 if dtype == 'float64' else torch.float32
    xs = torch.linspace(0, 1, n + 1, dtype=tdt)
    ys = torch.linspace(0, 1, ny + 1, dtype=tdt)
    f = -2 * np.pi ** 2 * torch.sin(np.pi * xs)[:, None] * torch.sin(np.pi * ys)[None, :]
    eq = api.Equation()
    eq.add({
        'd2u/dx2': {'coeff': 1, 'term': [0, 0], 'pow': 1},
        'd2u/dy2': {'coeff': 1, 'term': [1, 1], 'pow': 1},
        '-f': {'coeff': -f, 'term': [None], 'pow': 0},
    })
    return Problem(f'poisson_mat_p{derivative_points}', dom, bc, eq, 'mat', [],
                   dict(lambda_operator=1, lambda_bound=100, derivative_points=derivative_points),
                   mat_shape=(1, n + 1, ny + 1))


def poisson_robin_mat(api, dtype='float32', n=24, ny=31, derivative_points=2):
    """Poisson in mat mode with `robin` conditions on two edges (alpha * u + beta * bop, tedeous/eval.py:357-388 - the
    alpha term is counted again inside every beta term, quirk q5)."""
    dom = api.Domain()
    dom.variable('x', [0, 1], n, dtype=dtype)
    dom.variable('y', [0, 1], ny, dtype=dtype)

    def bop(func_coeff, deriv_coeff, axis):
        return {'u': {'coeff': func_coeff, 'term': [None], 'pow': 1},
                'du/dn': {'coeff': deriv_coeff, 'term': [axis], 'pow': 1}}
    bc = api.Conditions()
    bc.robin({'x': 0, 'y': [0, 1]}, operator=bop(1, -1, 0), value=lambda g: -g[:, 1])
    bc.robin({'x': 1, 'y': [0, 1]}, operator=bop(2, 0.5, 0), value=0.25)
    bc.dirichlet({'x': [0, 1], 'y': 0}, value=0)
    bc.dirichlet({'x': [0, 1], 'y': 1}, value=lambda g: torch.sin(np.pi * g[:, 0]))
    tdt = torch.float64 if dtype == 'float64' else torch.float32
    xs = torch.linspace(0, 1, n + 1, dtype=tdt)
    ys = torch.linspace(0, 1, ny + 1, dtype=tdt)
    f = -2 * np.pi ** 2 * torch.sin(np.pi * xs)[:, None] * torch.sin(np.pi * ys)[None, :]
    eq = api.Equation()
    eq.add({
        'd2u/dx2': {'coeff': 1, 'term': [0, 0], 'pow': 1},
        'd2u/dy2': {'coeff': 1, 'term': [1, 1], 'pow': 1},
        '-f': {'coeff': -f, 'term': [None], 'pow': 0},
    })
    return Problem('poisson_robin_mat', dom, bc, eq, 'mat', [],
                   dict(lambda_operator=1, lambda_bound=[10., 100.], derivative_points=derivative_points),
                   mat_shape=(1, n + 1, ny + 1))


def heat_mat(api, dtype='float32', n=32, nt=32, derivative_points=2):
    """Linear constant-coefficient operator with different stencil reach per axis: u_t - 0.1 u_xx + 0.5 u - f."""
    dom = api.Domain()
    dom.variable('x', [0, 1], n, dtype=dtype)
    dom.variable('t', [0, 1], nt, dtype=dtype)
    bc = api.Conditions()
    bc.dirichlet({'x': [0, 1], 't': 0}, value=lambda g: torch.sin(np.pi * g[:, 0]))
    bc.dirichlet({'x': 0, 't': [0, 1]}, value=0)
    bc.dirichlet({'x': 1, 't': [0, 1]}, value=0)
    eq = api.Equation()
    eq.add({
        'du/dt': {'coeff': 1, 'term': [1], 'pow': 1},
        '-a*d2u/dx2': {'coeff': -0.1, 'term': [0, 0], 'pow': 1},
        '0.5u': {'coeff': 0.5, 'term': [None], 'pow': 1},
        '-f': {'coeff': lambda g: -torch.sin(np.pi * g[0]) * torch.exp(-g[1]), 'term': [None], 'pow': 0},
        'c': {'coeff': 0.25, 'term': [None], 'pow': 0},
    })
    return Problem(f'heat_mat_p{derivative_points}', dom, bc, eq, 'mat', [],
                   dict(lambda_operator=1, lambda_bound=10, derivative_points=derivative_points),
                   mat_shape=(1, n + 1, nt + 1))


def heat_mat_causal(api, dtype='float32', n=24, nt=20, tol=0.02):
    """Causal loss in mat mode (tedeous/losses.py:137-182, n_t = grid.shape[1]: grid axis 0 is time): u_t - 0.1 u_xx + u^2
    = f with a tensor coefficient, time as the FIRST variable."""
    dom = api.Domain()
    dom.variable('t', [0, 1], nt, dtype=dtype)
    dom.variable('x', [0, 1], n, dtype=dtype)
    bc = api.Conditions()
    bc.dirichlet({'t': 0, 'x': [0, 1]}, value=lambda g: torch.sin(np.pi * g[:, 1]))
    bc.dirichlet({'t': [0, 1], 'x': 0}, value=0)
    bc.dirichlet({'t': [0, 1], 'x': 1}, value=0)
    tdt = torch.float64 if dtype == 'float64' else torch.float32
    c = 0.5 + torch.linspace(0, 1, (nt + 1) * (n + 1), dtype=tdt).reshape(nt + 1, n + 1)
    eq = api.Equation()
    eq.add({
        'du/dt': {'coeff': 1, 'term': [0], 'pow': 1},
        '-a*d2u/dx2': {'coeff': -0.1, 'term': [1, 1], 'pow': 1},
        'c(t,x)*u^2': {'coeff': c, 'term': [None], 'pow': 2},
        '-f': {'coeff': lambda g: -torch.sin(np.pi * g[1]) * torch.exp(-g[0]), 'term': [None], 'pow': 0},
    })
    return Problem('heat_mat_causal', dom, bc, eq, 'mat', [],
                   dict(lambda_operator=1, lambda_bound=10, derivative_points=2, tol=tol),
                   mat_shape=(1, nt + 1, n + 1))


def schrodinger_mat(api, dtype='float32', n=20, nt=28, derivative_points=2):
    """Two coupled fields on a 2-D grid in mat mode: u_t + 0.5 v_xx + (u^2 + v^2) v = 0, v_t - 0.5 u_xx - (u^2 + v^2) u = 0
    (examples/examples_schrodinger/example_schrodinger_matrix.py pattern: products of powers across fields)."""
    dom = api.Domain()
    dom.variable('x', [-2, 2], n, dtype=dtype)
    dom.variable('t', [0, 1], nt, dtype=dtype)
    bc = api.Conditions()
    bc.dirichlet({'x': [-2, 2], 't': 0}, value=lambda g: 2 / torch.cosh(g[:, 0]), var=0)
    bc.dirichlet({'x': [-2, 2], 't': 0}, value=0., var=1)
    bc.periodic([{'x': -2, 't': [0, 1]}, {'x': 2, 't': [0, 1]}], var=0)
    bc.periodic([{'x': -2, 't': [0, 1]}, {'x': 2, 't': [0, 1]}], var=1)
    eq = api.Equation()
    eq.add({'du/dt': {'coeff': 1, 'term': [1], 'pow': 1, 'var': 0},
            '1/2*d2v/dx2': {'coeff': 0.5, 'term': [0, 0], 'pow': 1, 'var': 1},
            'v*u**2': {'coeff': 1, 'term': [[None], [None]], 'pow': [1, 2], 'var': [1, 0]},
            'v**3': {'coeff': 1, 'term': [None], 'pow': 3, 'var': 1}})
    eq.add({'dv/dt': {'coeff': 1, 'term': [1], 'pow': 1, 'var': 1},
            '-1/2*d2u/dx2': {'coeff': -0.5, 'term': [0, 0], 'pow': 1, 'var': 0},
            '-u*v**2': {'coeff': -1, 'term': [[None], [None]], 'pow': [1, 2], 'var': [0, 1]},
            '-u**3': {'coeff': -1, 'term': [None], 'pow': 3, 'var': 0}})
    return Problem('schrodinger_mat', dom, bc, eq, 'mat', [],
                   dict(lambda_operator=1, lambda_bound=[5., 7.], derivative_points=derivative_points),
                   mat_shape=(2, n + 1, nt + 1))


def kdv_mat(api, dtype='float32', n=24, derivative_points=2):
    """Nonlinear mat-mode operator with 3rd derivative, periodic + operator conditions
    (examples/examples_korteweg_de_vries/example_KdV_matrix.py pattern)."""
    dom = api.Domain()
    dom.variable('x', [0, 1], n, dtype=dtype)
    dom.variable('t', [0, 1], n, dtype=dtype)
    bc = api.Conditions()
    bc.dirichlet({'x': [0, 1], 't': 0}, value=lambda g: torch.sin(2 * np.pi * g[:, 0]))
    bc.periodic([{'x': 0, 't': [0, 1]}, {'x': 1, 't': [0, 1]}])
    bc.operator({'x': 1, 't': [0, 1]}, operator={'du/dx': {'coeff': 1, 'term': [0], 'pow': 1}}, value=0.5)
    eq = api.Equation()
    eq.add({
        'du/dt': {'coeff': 1, 'term': [1], 'pow': 1},
        '6u*du/dx': {'coeff': 6, 'term': [[None], [0]], 'pow': [1, 1]},
        'd3u/dx3': {'coeff': 1, 'term': [0, 0, 0], 'pow': 1},
        '-f': {'coeff': lambda g: -torch.sin(g[0]) * torch.cos(g[1]), 'term': [None], 'pow': 0},
    })
    return Problem(f'kdv_mat_p{derivative_points}', dom, bc, eq, 'mat', [],
                   dict(lambda_operator=1, lambda_bound=[10., 20., 30.], derivative_points=derivative_points),
                   mat_shape=(1, n + 1, n + 1))


# --- weak form (examples/examples_lotka_volterra/example_weak_LotkaVolterra.py:26-121) ---------------------
def lotka_weak(api, dtype='float32', n=40, mode='NN', layers=(1, 32, 32, 2)):
    alpha, beta, delta, gamma, x0, y0, tmax = 20., 20., 20., 20., 4., 2., 1.
    dom = api.Domain()
    dom.variable('t', [0, tmax], n, dtype=dtype)
    h = tmax / n
    bc = api.Conditions()
    bc.dirichlet({'t': 0}, value=x0, var=0)
    bc.dirichlet({'t': 0}, value=y0, var=1)
    eq = api.Equation()
    eq.add({'dx/dt': {'coeff': 1, 'term': [0], 'pow': 1, 'var': [0]},
            '-x*alpha': {'coeff': -alpha, 'term': [None], 'pow': 1, 'var': [0]},
            '+beta*x*y': {'coeff': beta, 'term': [[None], [None]], 'pow': [1, 1], 'var': [0, 1]}})
    eq.add({'dy/dt': {'coeff': 1, 'term': [0], 'pow': 1, 'var': [1]},
            '+y*delta': {'coeff': delta, 'term': [None], 'pow': 1, 'var': [1]},
            '-gamma*x*y': {'coeff': -gamma, 'term': [[None], [None]], 'pow': [1, 1], 'var': [0, 1]}})

    if mode == 'mat':           # examples/examples_lotka_volterra/example_LV_mat.py: 1-D grid, two fields
        return Problem('lotka_mat', dom, bc, eq, 'mat', [], dict(lambda_operator=1, lambda_bound=100, derivative_points=2),
                       mat_shape=(2, n + 1))

    def v(grid):
        return (0.5 + 0.5 * torch.sin(grid[:, 0])) * (2 / h) ** 0.5 / 10
    kw = dict(lambda_operator=1, lambda_bound=100, weak_form=[v])
    if mode == 'NN':
        kw['h'] = h
    return Problem(f'lotka_weak_{mode}', dom, bc, eq, mode, list(layers), kw)


def wave_weak(api, dtype='float32', n=16):
    """2-D grid: two nested integration passes (eval.py:13-52 squares the integrand in each of them)."""
    prob = wave(api, dtype, n=n, mode='autograd', layers=(2, 32, 32, 1))
    prob.compile_kwargs = dict(prob.compile_kwargs)
    prob.compile_kwargs['weak_form'] = [lambda grid: 0.5 + 0.25 * torch.sin(3 * grid[:, 0]) * torch.cos(2 * grid[:, 1])]
    prob.name = 'wave_weak_autograd'
    return prob


# --- Robin conditions (examples/examples_poisson/example_poisson_2d_many_subdomains.py:46-75) ----------------
def poisson_robin(api, dtype='float32', n=20, layers=(2, 32, 32, 1)):
    """u_xx + u_yy = f on the unit square; u -/+ du/dn = g on the four edges as `robin` conditions (alpha * u +
    beta * bop, the reference's quirk q5 included: tedeous/eval.py:357-388), autograd mode."""
    dom = api.Domain()
    dom.variable('x', [0, 1], n, dtype=dtype)
    dom.variable('y', [0, 1], n, dtype=dtype)

    def bop(func_coeff, deriv_coeff, axis):
        return {'u': {'coeff': func_coeff, 'term': [None], 'pow': 1},
                'du/dn': {'coeff': deriv_coeff, 'term': [axis], 'pow': 1}}
    bc = api.Conditions()
    bc.robin({'x': 0, 'y': [0, 1]}, operator=bop(1, -1, 0), value=lambda g: -g[:, 1])
    bc.robin({'x': 1, 'y': [0, 1]}, operator=bop(1, 1, 0), value=lambda g: -g[:, 1])
    bc.robin({'x': [0, 1], 'y': 0}, operator=bop(2, -0.5, 1), value=lambda g: torch.sin(g[:, 0]))
    bc.robin({'x': [0, 1], 'y': 1}, operator=bop(1, 1, 1), value=0.25)
    eq = api.Equation()
    eq.add({
        'd2u/dx2': {'coeff': 1, 'term': [0, 0], 'pow': 1},
        'd2u/dy2': {'coeff': 1, 'term': [1, 1], 'pow': 1},
        '-f': {'coeff': lambda g: -torch.sin(np.pi * g[:, 0]) * torch.cos(np.pi * g[:, 1]), 'term': [None], 'pow': 0},
    })
    return Problem('poisson_robin_autograd', dom, bc, eq, 'autograd', list(layers), dict(lambda_operator=1, lambda_bound=10),
                   init='xavier_b')


# --- mixed partial derivatives (the reference differentiates along any axis list, tedeous/derivative.py:92-97) ---------
def mixed_elliptic(api, dtype='float32', n=20, mode='autograd', layers=(2, 100, 100, 100, 1), h=0.01):
    """u_xx + 1.5 u_xy + 2 u_yy - 0.5 u u_yx = f with a mixed-partial boundary operator: second-order mixed partials
    next to the pure ones (directions x, y, x + y: the tensor-core signature (2, 2, 2))."""
    dom = api.Domain()
    dom.variable('x', [0, 1], n, dtype=dtype)
    dom.variable('y', [0, 1], n, dtype=dtype)
    bc = api.Conditions()
    bc.dirichlet({'x': 0, 'y': [0, 1]}, value=lambda g: torch.sin(g[:, 1]))
    bc.dirichlet({'x': 1, 'y': [0, 1]}, value=lambda g: torch.cos(g[:, 1]))
    bc.dirichlet({'x': [0, 1], 'y': 0}, value=lambda g: g[:, 0])
    if mode == 'autograd':
        bc.operator({'x': [0, 1], 'y': 1}, operator={'d2u/dxdy': {'coeff': 1, 'term': [0, 1], 'pow': 1, 'var': 0},
                                                     'u': {'coeff': 0.5, 'term': [None], 'pow': 1, 'var': 0}}, value=0.2)
    else:
        bc.dirichlet({'x': [0, 1], 'y': 1}, value=0.2)
    eq = api.Equation()
    eq.add({
        'd2u/dx2': {'coeff': 1, 'term': [0, 0], 'pow': 1, 'var': 0},
        'd2u/dxdy': {'coeff': 1.5, 'term': [0, 1], 'pow': 1, 'var': 0},
        'd2u/dy2': {'coeff': 2, 'term': [1, 1], 'pow': 1, 'var': 0},
        'u*d2u/dydx': {'coeff': -0.5, 'term': [[None], [1, 0]], 'pow': [1, 1], 'var': [0, 0]},
        '-f': {'coeff': lambda g: -torch.sin(np.pi * g[:, 0]) * g[:, 1], 'term': [None], 'pow': 0},
    })
    kw = dict(lambda_operator=1, lambda_bound=10)
    if mode == 'NN':
        kw['h'] = h
    return Problem(f'mixed_elliptic_{mode}', dom, bc, eq, mode, list(layers), kw, init='xavier_b')


def mixed_bbm(api, dtype='float32', n=16, layers=(2, 32, 32, 1)):
    """Benjamin-Bona-Mahony type equation u_t + u u_x - 0.05 u_xxt = 0: a THIRD-order mixed partial, and a second-order
    mixed partial alone (u_xt, directions x + t and x - t), autograd mode."""
    dom = api.Domain()
    dom.variable('x', [0, 1], n, dtype=dtype)
    dom.variable('t', [0, 1], n, dtype=dtype)
    bc = api.Conditions()
    bc.dirichlet({'x': [0, 1], 't': 0}, value=lambda g: torch.sin(np.pi * g[:, 0]))
    bc.dirichlet({'x': 0, 't': [0, 1]}, value=0.)
    bc.dirichlet({'x': 1, 't': [0, 1]}, value=0.)
    eq = api.Equation()
    eq.add({
        'du/dt': {'coeff': 1, 'term': [1], 'pow': 1, 'var': 0},
        'u*du/dx': {'coeff': 1, 'term': [[None], [0]], 'pow': [1, 1], 'var': [0, 0]},
        '-0.05*d3u/dx2dt': {'coeff': -0.05, 'term': [0, 0, 1], 'pow': 1, 'var': 0},
        '0.1*d2u/dxdt': {'coeff': 0.1, 'term': [0, 1], 'pow': 1, 'var': 0},
    })
    return Problem('mixed_bbm_autograd', dom, bc, eq, 'autograd', list(layers), dict(lambda_operator=1, lambda_bound=10),
                   init='xavier_b')


def monge_ampere(api, dtype='float32', n=20, layers=(2, 64, 64, 64, 1)):
    """Monge-Ampere equation u_xx u_yy - (u_xy)^2 = f: a SQUARED mixed partial (the power of the polarisation sum is
    expanded into products of directional derivatives), autograd mode."""
    dom = api.Domain()
    dom.variable('x', [0, 1], n, dtype=dtype)
    dom.variable('y', [0, 1], n, dtype=dtype)
    bc = api.Conditions()
    exact = lambda g: torch.exp(0.5 * (g[:, 0] ** 2 + g[:, 1] ** 2))
    bc.dirichlet({'x': 0, 'y': [0, 1]}, value=exact)
    bc.dirichlet({'x': 1, 'y': [0, 1]}, value=exact)
    bc.dirichlet({'x': [0, 1], 'y': 0}, value=exact)
    bc.dirichlet({'x': [0, 1], 'y': 1}, value=exact)
    eq = api.Equation()
    eq.add({
        'uxx*uyy': {'coeff': 1, 'term': [[0, 0], [1, 1]], 'pow': [1, 1], 'var': [0, 0]},
        '-uxy^2': {'coeff': -1, 'term': [0, 1], 'pow': 2, 'var': 0},
        '-f': {'coeff': lambda g: -(1 + g[:, 0] ** 2 + g[:, 1] ** 2) * torch.exp(g[:, 0] ** 2 + g[:, 1] ** 2),
               'term': [None], 'pow': 0},
    })
    return Problem('monge_ampere_autograd', dom, bc, eq, 'autograd', list(layers), dict(lambda_operator=1, lambda_bound=10),
                   init='xavier_b')


# --- callable 'pow' (examples/examples_heat/example_heat_2d_long_time.py:94-99) ------------------------------------------
def heat_callable_pow(api, dtype='float32', n=20, mode='autograd', layers=(2, 32, 32, 1), h=0.01):
    """u_t - 0.05 u_xx + c(x, t) sin(3 u^2) + 0.3 tanh(u) u_x^2 = 0: a callable power on the value (the shipped example's
    source term) and a chain of a callable and a numeric power over two factors (`der = pow_j(der * factor_j)`,
    tedeous/derivative.py:52-55, 126-129)."""
    dom = api.Domain()
    dom.variable('x', [0, 1], n, dtype=dtype)
    dom.variable('t', [0, 1], n, dtype=dtype)
    bc = api.Conditions()
    bc.dirichlet({'x': [0, 1], 't': 0}, value=lambda g: torch.sin(np.pi * g[:, 0]))
    bc.dirichlet({'x': 0, 't': [0, 1]}, value=0.)
    bc.dirichlet({'x': 1, 't': [0, 1]}, value=0.)
    eq = api.Equation()
    eq.add({
        'du/dt': {'coeff': 1, 'term': [1], 'pow': 1, 'var': 0},
        '-0.05*d2u/dx2': {'coeff': -0.05, 'term': [0, 0], 'pow': 1, 'var': 0},
        'c*sin(3u^2)': {'coeff': lambda g: 1. + g[:, 0] * g[:, 1], 'term': [None], 'pow': lambda u: torch.sin(3. * u ** 2),
                        'var': 0},
        '0.3*tanh(u)*(du/dx)^2': {'coeff': 0.3, 'term': [[None], [0]], 'pow': [lambda z: torch.tanh(z), 2], 'var': [0, 0]},
    })
    kw = dict(lambda_operator=1, lambda_bound=10)
    if mode == 'NN':
        kw['h'] = h
    return Problem(f'heat_callable_{mode}', dom, bc, eq, mode, list(layers), kw, init='xavier_b')


def trained(prob: Problem, steps: int) -> Problem:
    prob.train_steps = steps
    return prob


# fixtures too large for the every-fixture CPU oracle sweep (tests/test_oracle_golden.py checks them in fp64 only)
LARGE = ('wave_autograd_1e5', 'kdv_autograd_1e5', 'wave_autograd_3e5')

ZOO: Dict[str, Callable] = {
    'poisson_robin_autograd': lambda api, dt: poisson_robin(api, dt),
    'mixed_elliptic_autograd': lambda api, dt: mixed_elliptic(api, dt, mode='autograd'),
    'mixed_elliptic_NN': lambda api, dt: mixed_elliptic(api, dt, n=16, mode='NN', layers=(2, 32, 32, 1)),
    'mixed_bbm_autograd': lambda api, dt: mixed_bbm(api, dt),
    'monge_ampere_autograd': lambda api, dt: monge_ampere(api, dt),
    'burgers_inverse_autograd': lambda api, dt: burgers_inverse(api, dt, mode='autograd'),
    'heat_callable_autograd': lambda api, dt: heat_callable_pow(api, dt, mode='autograd'),
    'heat_callable_NN': lambda api, dt: heat_callable_pow(api, dt, n=16, mode='NN'),
    # ~10^5 points: the sizes at which the tensor-core kernels are chosen automatically (impl = 0)
    'wave_autograd_1e5': lambda api, dt: wave(api, dt, n=315, mode='autograd'),
    'kdv_autograd_1e5': lambda api, dt: kdv(api, dt, nx=399, nt=249, mode='autograd'),
    'wave_autograd_3e5': lambda api, dt: wave(api, dt, n=547, mode='autograd'),
    # trained-ish states: 200 Adam steps of the reference (residual cancellation is worst near convergence)
    'burgers_autograd_trained': lambda api, dt: trained(burgers(api, dt, n=40, mode='autograd'), 200),
    'wave_autograd_trained': lambda api, dt: trained(wave(api, dt, n=40, mode='autograd'), 200),
    'lotka_weak_NN': lambda api, dt: lotka_weak(api, dt, mode='NN'),
    'lotka_weak_autograd': lambda api, dt: lotka_weak(api, dt, mode='autograd'),
    'wave_weak_autograd': lambda api, dt: wave_weak(api, dt),
    # name -> (builder, kwargs).  Sizes chosen so the reference itself preprocesses them in seconds.
    'burgers_NN_cfg1': lambda api, dt: burgers(api, dt, n=100, mode='NN'),
    'burgers_NN_small': lambda api, dt: burgers(api, dt, n=24, mode='NN', layers=(2, 32, 32, 1)),
    'burgers_autograd_4h': lambda api, dt: burgers(api, dt, n=40, mode='autograd', layers=(2, 100, 100, 100, 100, 1)),
    'burgers_autograd_causal': lambda api, dt: burgers(api, dt, n=30, mode='autograd', layers=(2, 32, 32, 1), tol=2.0),
    'burgers_NN_causal': lambda api, dt: burgers(api, dt, n=20, mode='NN', layers=(2, 32, 32, 1), tol=2.0),
    'wave_autograd': lambda api, dt: wave(api, dt, n=40, mode='autograd'),
    'wave_NN': lambda api, dt: wave(api, dt, n=20, mode='NN', layers=(2, 32, 32, 1)),
    'kdv_autograd': lambda api, dt: kdv(api, dt, nx=30, nt=30, mode='autograd'),
    'kdv_NN': lambda api, dt: kdv(api, dt, nx=20, nt=20, mode='NN', layers=(2, 32, 32, 1)),
    'navier_stokes_autograd': lambda api, dt: navier_stokes(api, dt, n=8),
    'legendre_autograd': lambda api, dt: legendre_ode(api, dt, mode='autograd'),
    'legendre_NN': lambda api, dt: legendre_ode(api, dt, mode='NN'),
    'legendre_mat_1d': lambda api, dt: legendre_ode(api, dt, n=48, mode='mat'),
    'lotka_mat_1d': lambda api, dt: lotka_weak(api, dt, n=64, mode='mat'),
    'nonlinear_mix_autograd': lambda api, dt: nonlinear_mix(api, dt, mode='autograd'),
    'nonlinear_mix_NN': lambda api, dt: nonlinear_mix(api, dt, mode='NN'),
    'poisson_mat_p2': lambda api, dt: poisson_mat(api, dt, n=32, derivative_points=2),
    'poisson_robin_mat': lambda api, dt: poisson_robin_mat(api, dt),
    'heat_mat_causal': lambda api, dt: heat_mat_causal(api, dt),
    'poisson_mat_p3': lambda api, dt: poisson_mat(api, dt, n=24, derivative_points=3),
    'kdv_mat_p2': lambda api, dt: kdv_mat(api, dt, n=24, derivative_points=2),
    'schrodinger_mat_p2': lambda api, dt: schrodinger_mat(api, dt),
    # grids with n1 % 4 == 0: served by the vectorised cross-stencil kernel (one tile, every cell next to an edge)
    'poisson_mat_p2_rect': lambda api, dt: poisson_mat(api, dt, n=40, ny=63, derivative_points=2),
    'poisson_mat_p3_rect': lambda api, dt: poisson_mat(api, dt, n=24, ny=43, derivative_points=3),
    'heat_mat_p2': lambda api, dt: heat_mat(api, dt, n=31, nt=47, derivative_points=2),
}
