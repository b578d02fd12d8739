"""Shared helpers for the test-suite: load goldens, build problems through this repo's front end, run the
oracle on them."""
import os

import numpy as np
import torch

import problems
import torch_de_solver_b200 as tdb
from oracle import tedeous_oracle as orc

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_golden(name, dtype):
    return dict(np.load(os.path.join(GOLDEN_DIR, f'{name}.{dtype}.npz'), allow_pickle=False))


def set_weights(params, flat):
    off = 0
    with torch.no_grad():
        for p in params:
            n = p.numel()
            p.copy_(torch.as_tensor(flat[off:off + n]).reshape(p.shape).to(p.dtype))
            off += n
    assert off == len(flat)


def build(name, dtype, weights=None):
    """-> (problem, grid, bconds, model) on the CPU through this repo's Domain / Conditions / Equation."""
    tdt = torch.float64 if dtype == 'float64' else torch.float32
    prob = problems.ZOO[name](tdb, dtype)
    grid = prob.domain.build(prob.mode)
    bconds = prob.conditions.build(prob.domain.variable_dict)
    if prob.mode == 'mat':
        model = problems.make_mat_model(prob.mat_shape, tdt)
        if weights is not None:
            model = torch.as_tensor(weights).reshape(prob.mat_shape).to(tdt)
    else:
        model = problems.make_net(prob.net_layers, tdt, prob.init)
        if weights is not None:
            set_weights(list(model.parameters()), weights)
    return prob, grid, bconds, model


def oracle_solution(prob, grid, bconds, model):
    kw = prob.compile_kwargs
    return orc.OracleSolution(grid, prob.equation.equation_lst, bconds, model, prob.mode,
                              kw['lambda_operator'], kw['lambda_bound'], h=kw.get('h', 0.001),
                              derivative_points=kw.get('derivative_points', 2), tol=kw.get('tol', 0),
                              weak_form=kw.get('weak_form'))


def oracle_eval(name, dtype, weights=None):
    prob, grid, bconds, model = build(name, dtype, weights)
    sol = oracle_solution(prob, grid, bconds, model)
    params = [model.requires_grad_()] if prob.mode == 'mat' else list(model.parameters())
    loss, loss_n, grads = orc.loss_and_grad(sol, params)
    flat = torch.cat([g.reshape(-1) for g in grads]).double().numpy()
    return sol, float(loss), float(loss_n), flat
