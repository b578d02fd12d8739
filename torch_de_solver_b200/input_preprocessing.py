"""Setup-time normalisation of equations and conditions (tedeous/input_preprocessing.py).

`lambda_prepare` and `EquationMixin.equation_unify` keep the reference's exact semantics.  The three
`Equation_*` classes keep the reference's constructor signatures and `operator_prepare()/bnd_prepare()`
methods, but they return the *unified* term dicts only - no shifted grid copies (NN) and no O(n_bnd*N)
isclose scans (mat): `plan.lower_problem` turns them into the flat term table / jet spec / segment
table the CUDA kernels consume."""
from copy import deepcopy
from typing import Callable, Union
import torch

from .device import check_device
from .points_type import Points_type


def lambda_prepare(val: torch.Tensor, lambda_: Union[int, float, list, torch.Tensor],
                   dtype: torch.dtype = None) -> torch.Tensor:
    """Scalar / list -> [1, n_cols] tensor of lambdas (tedeous/input_preprocessing.py:14-43)."""
    if isinstance(lambda_, torch.Tensor):
        return lambda_
    if isinstance(lambda_, (int, float)):
        try:
            lambdas = torch.ones(val.shape[-1]) * lambda_
        except Exception:
            lambdas = torch.tensor(lambda_)
    elif isinstance(lambda_, list):
        lambdas = torch.tensor(lambda_)
    else:
        raise TypeError(f'bad lambda type {type(lambda_)}')
    if dtype:
        lambdas = lambdas.to(dtype)
    return lambdas.reshape(1, -1)


class EquationMixin:
    @staticmethod
    def equation_unify(equation: dict) -> dict:
        """List-ify 'pow' / 'var' / the derivative spec of every term, add 'var' = 0 when missing
        (tedeous/input_preprocessing.py:52-81).  The derivative spec is the term's 2nd key."""
        for label in equation.keys():
            term = equation[label]
            dif_dir = list(term.keys())[1]
            scalar_pow = isinstance(term['pow'], (int, float)) or callable(term['pow'])
            if 'var' not in term:
                if scalar_pow:
                    term[dif_dir] = [term[dif_dir]]
                    term['pow'] = [term['pow']]
                    term['var'] = [0]
                elif isinstance(term['pow'], list):
                    term['var'] = [0 for _ in term['pow']]
                continue
            if scalar_pow:
                term[dif_dir] = [term[dif_dir]]
                term['pow'] = [term['pow']]
                term['var'] = [term['var']]
        return equation


def _as_equation_list(operator) -> list:
    if isinstance(operator, list) and isinstance(operator[0], dict):
        return operator
    return [operator]


def _check_coeff(coeff):
    if isinstance(coeff, (int, float)) or callable(coeff):
        return coeff
    if isinstance(coeff, torch.nn.parameter.Parameter):
        return coeff
    if isinstance(coeff, torch.Tensor):
        return check_device(coeff)
    raise NameError('"coeff" should be: torch.Tensor or callable or int or float!')


class _EquationBase(EquationMixin):
    mode = None

    def __init__(self, grid, operator, bconds):
        self.grid = grid
        self.operator = operator
        self.bconds = bconds

    def operator_prepare(self) -> list:
        prepared = []
        for eq in _as_equation_list(self.operator):
            eq = self.equation_unify(eq)
            for label in eq:
                eq[label]['coeff'] = _check_coeff(eq[label]['coeff'])
            prepared.append(eq)
        return prepared

    def bnd_prepare(self):
        if self.bconds is None:
            return None
        for bcond in self.bconds:
            if bcond['bop'] is not None:
                bcond['bop'] = self.equation_unify(bcond['bop'])
        return self.bconds


class Equation_NN(_EquationBase, Points_type):
    """tedeous/input_preprocessing.py:173-408 (constructor signature kept)."""
    mode = 'NN'

    def __init__(self, grid, operator, bconds, h: float = 0.001, inner_order: str = '1',
                 boundary_order: str = '2'):
        Points_type.__init__(self, grid)
        _EquationBase.__init__(self, grid, operator, bconds)
        self.h = h
        self.inner_order = inner_order
        self.boundary_order = boundary_order


class Equation_autograd(_EquationBase):
    """tedeous/input_preprocessing.py:411-508."""
    mode = 'autograd'


class Equation_mat(_EquationBase):
    """tedeous/input_preprocessing.py:511-594."""
    mode = 'mat'


class Operator_bcond_preproc:
    """set_strategy(mode) -> Equation_{NN, autograd, mat}  (tedeous/input_preprocessing.py:597-644)."""

    def __init__(self, grid, operator, bconds, h: float = 0.001, inner_order: str = '1',
                 boundary_order: str = '2'):
        self.grid = check_device(grid)
        self.operator = operator
        self.bconds = bconds
        self.h = h
        self.inner_order = inner_order
        self.boundary_order = boundary_order

    def set_strategy(self, strategy: str):
        if strategy == 'NN':
            return Equation_NN(self.grid, self.operator, self.bconds, h=self.h,
                               inner_order=self.inner_order, boundary_order=self.boundary_order)
        if strategy == 'mat':
            return Equation_mat(self.grid, self.operator, self.bconds)
        if strategy == 'autograd':
            return Equation_autograd(self.grid, self.operator, self.bconds)
        raise ValueError(f'unknown mode {strategy!r}')
