"""Compatibility views with the reference's `Operator` / `Bounds` interface (tedeous/eval.py:90-232, 235-461).

Inside `Solution` they are thin views over the fused plan: `operator_compute()` / `apply_bcs()` launch the
forward-only kernel (`tdb200_eval_fields`) and return the per-point fields.  Constructed stand-alone with
the reference's signatures they build a small plan of their own, so code that used the reference's seams
directly (optimizers/closure.py:113-114, landscape_visualization/_aux/PINN_loss_data.py:21-27) keeps working.
In modes NN / autograd the returned tensors are differentiable w.r.t. the network parameters when autograd is
recording (backward = one fused launch in vector-Jacobian mode, `tdb200_plan_set_field_seeds`); in mat mode they are
detached."""
from typing import List, Tuple

import torch

from .device import check_device


def _solution_for(grid, prepared_operator, prepared_bconds, model, mode, derivative_points):
    from .input_preprocessing import Operator_bcond_preproc
    from .solution import Solution
    eq = Operator_bcond_preproc(grid, prepared_operator, prepared_bconds).set_strategy(mode)
    return Solution(grid, eq, model, mode, None, 1, 1, derivative_points=derivative_points)


def _fields_of(sol):
    """(op, bval, true_bval): attached to the autograd graph when that is possible and asked for."""
    sol._fields_cache = None
    if sol.mode != 'mat' and torch.is_grad_enabled() and any(p.requires_grad for p in sol.model.parameters()):
        return sol._fields_differentiable()
    return sol._fields()


class Operator:
    def __init__(self, grid, prepared_operator, model, mode, weak_form=None, derivative_points=2,
                 batch_size=None):
        grid = check_device(grid)
        self.weak_form = weak_form
        # a plan needs at least one condition; a single dummy Dirichlet row is never read back
        dummy = [{'bnd': (grid[:1] if mode != 'mat' else grid.reshape(grid.shape[0], -1)[:, :1].T),
                  'bop': None, 'bval': torch.zeros(1, device=grid.device), 'var': 0, 'type': 'dirichlet'}]
        self._sol = _solution_for(grid, prepared_operator, dummy, model, mode, derivative_points)
        self._init_common()

    def _init_common(self):
        if not hasattr(self, 'weak_form'):
            self.weak_form = self._sol.weak_form
        self.grid = self._sol.grid
        self.model = self._sol.model
        self.mode = self._sol.mode
        self.batch_size = getattr(self._sol, 'batch_size', None)
        self.n_batches = getattr(self._sol, 'n_batches', 1)
        self.current_batch_i = 0

    @classmethod
    def _from_solution(cls, sol):
        self = cls.__new__(cls)
        self._sol = sol
        self._init_common()
        return self

    def _pde_compute(self) -> torch.Tensor:
        return _fields_of(self._sol)[0]

    def _weak_pde_compute(self) -> torch.Tensor:
        from .losses import weak_operator
        return weak_operator(self._pde_compute(), self._sol._ir.interior_points, self.weak_form)

    def operator_compute(self) -> torch.Tensor:
        if self.weak_form in (None, []):
            return self._pde_compute()
        return self._weak_pde_compute()


class Bounds:
    def __init__(self, grid, prepared_bconds, model, mode, weak_form=None, derivative_points=2):
        grid = check_device(grid)
        dummy_op = [{'u': {'coeff': 1., 'u': [None], 'pow': 1, 'var': 0}}]
        self._sol = _solution_for(grid, dummy_op, prepared_bconds, model, mode, derivative_points)
        self.grid, self.model, self.mode = self._sol.grid, self._sol.model, mode

    @classmethod
    def _from_solution(cls, sol):
        self = cls.__new__(cls)
        self._sol = sol
        self.grid, self.model, self.mode = sol.grid, sol.model, sol.mode
        return self

    def apply_bcs(self) -> Tuple[torch.Tensor, torch.Tensor, List[str], List[int]]:
        _, bval, true_bval = _fields_of(self._sol)
        return bval, true_bval, list(self._sol.bval_keys), list(self._sol.bval_length)
