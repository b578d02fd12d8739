"""`Derivative(model, derivative_points).set_strategy(mode).take_derivative(term, grid_points)`
(tedeous/derivative.py:326-363) on top of the fused plan: the term is lowered to a one-term operator and
evaluated by the forward-only kernel.

Differences from the reference, by design: the term is the *unified* dict (`{'coeff', <dir>: [[axes]..],
'pow': [..], 'var': [..]}`), also in NN mode - interior derivatives are exact Taylor jets at `grid_points`,
not stencil sums over pre-shifted grid copies (derivative.py:47-51)."""
import torch

from .eval import Operator
from .input_preprocessing import EquationMixin


class _Strategy:
    def __init__(self, model, mode, derivative_points):
        self.model, self.mode, self.derivative_points = model, mode, derivative_points

    def take_derivative(self, term: dict, grid_points: torch.Tensor = None) -> torch.Tensor:
        if grid_points is None:
            raise ValueError('grid_points is required')
        op = EquationMixin.equation_unify({'term': dict(term)})
        mode = 'autograd' if self.mode == 'NN' else self.mode
        out = Operator(grid_points, [op], self.model, mode, None, self.derivative_points).operator_compute()
        if self.mode == 'mat':
            return out.reshape(self.model.shape)
        return out.reshape(-1, 1)


class Derivative_NN(_Strategy):
    def __init__(self, model):
        super().__init__(model, 'NN', 2)


class Derivative_autograd(_Strategy):
    def __init__(self, model):
        super().__init__(model, 'autograd', 2)


class Derivative_mat(_Strategy):
    def __init__(self, model, derivative_points):
        super().__init__(model, 'mat', derivative_points)


class Derivative:
    def __init__(self, model, derivative_points):
        self.model = model
        self.derivative_points = derivative_points

    def set_strategy(self, strategy: str):
        if strategy == 'NN':
            return Derivative_NN(self.model)
        if strategy == 'autograd':
            return Derivative_autograd(self.model)
        if strategy == 'mat':
            return Derivative_mat(self.model, self.derivative_points)
        raise ValueError(strategy)
