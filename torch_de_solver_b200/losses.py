"""`Losses(mode, weak_form, n_t, tol).compute(op, bval, true_bval, lambda_op, lambda_bound, save_graph)`
(tedeous/losses.py:230-263) for callers that hold per-point fields.  `Solution.evaluate` does NOT go through
this class - the fused kernel reduces the residuals itself; this is the same formula applied to tensors the
caller already has (a handful of device ops on [N, n_eq] / [max_len, n_types] arrays)."""
from typing import Tuple

import torch

from .input_preprocessing import lambda_prepare


class Losses:
    def __init__(self, mode, weak_form, n_t, tol, n_t_operation=None):
        if weak_form not in (None, []):
            raise NotImplementedError('weak-form loss is not implemented')
        self.mode, self.weak_form, self.n_t, self.tol, self.n_t_operation = mode, weak_form, n_t, tol, n_t_operation

    def _loss_bcs(self, bval, true_bval, lambda_bound):
        diff = torch.mean((bval - true_bval) ** 2, 0)
        return diff @ lambda_bound.T.to(diff), diff

    def _default_loss(self, operator, bval, true_bval, lambda_op, lambda_bound, save_graph=True):
        op = torch.mean(operator ** 2, 0)
        loss_bnd, diff = self._loss_bcs(bval, true_bval, lambda_bound)
        loss = op @ lambda_op.T.to(op) + loss_bnd
        with torch.no_grad():
            loss_normalized = op.sum().reshape(1) + diff.sum().reshape(1)
        return (loss if save_graph else loss.detach()), loss_normalized

    def _causal_loss(self, operator, bval, true_bval, lambda_op, lambda_bound):
        """losses.py:137-182 (grid column 0 is time; lambda_op is ignored - SURVEY B.1 q7)."""
        if self.n_t_operation is not None:
            self.n_t = self.n_t_operation(operator)
        res = torch.sum(operator ** 2, dim=1).reshape(self.n_t, -1)
        with torch.no_grad():
            w = torch.exp(-self.tol * (torch.cumsum(res, 0) - res))
        loss_oper = torch.mean(w * res)
        loss_bnd, diff = self._loss_bcs(bval, true_bval, lambda_bound)
        with torch.no_grad():
            loss_normalized = loss_oper + diff.sum()
        return loss_oper + loss_bnd, loss_normalized.reshape(1)

    def compute(self, operator, bval, true_bval, lambda_op, lambda_bound,
                save_graph: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
        if bval is None:
            raise ValueError('a problem without boundary conditions has no finite loss '
                             '(the reference prints a warning and returns inf, losses.py:250-255)')
        lambda_op = lambda_prepare(operator, lambda_op)
        lambda_bound = lambda_prepare(bval, lambda_bound)
        if self.tol != 0:
            return self._causal_loss(operator, bval, true_bval, lambda_op, lambda_bound)
        return self._default_loss(operator, bval, true_bval, lambda_op, lambda_bound, save_graph)
