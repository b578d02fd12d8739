"""`Losses(mode, weak_form, n_t, tol).compute(op, bval, true_bval, lambda_op, lambda_bound, save_graph)`
(tedeous/losses.py:230-263) for callers that hold per-point fields.  `Solution.evaluate` does NOT go through
this class for the default and causal losses - the fused kernel reduces the residuals itself; this is the same
formula applied to tensors the caller already has (a handful of device ops on [N, n_eq] / [max_len, n_types] arrays).
The weak-form loss (losses.py:184-228) is the exception: its nested integrals are not a sum over points, so
`Solution.evaluate` forms them here from the per-point fields of a forward launch and gets the parameter gradient
from the fused kernel in vector-Jacobian mode (`tdb200_plan_set_field_seeds`)."""
from typing import Tuple

import torch

from .input_preprocessing import lambda_prepare


def integration(func: torch.Tensor, grid: torch.Tensor, power: int = 2):
    """One integration pass of the weak form (tedeous/eval.py:13-52), without the reference's per-point Python loop
    and `.item()` host syncs: trapezoid rule along the LAST grid column inside runs of equal values of the column
    before it (a single run when the grid has one column), the integrand raised to `power` (the reference's default 2
    is applied by every pass).  -> (run integrals [n_runs] or a 0-d tensor, grid of the run starts without the last
    column or None)."""
    f = func ** power
    seg = (grid[1:, -1] - grid[:-1, -1]) * (f[1:] + f[:-1]) / 2          # trapezoids between consecutive rows
    if grid.shape[-1] == 1:
        return seg.sum(), None
    key = grid[:, -2]
    start = torch.ones(grid.shape[0], dtype=torch.bool, device=grid.device)
    start[1:] = key[1:] != key[:-1]                                       # a new run: the pair across it is skipped
    run = torch.cumsum(start, 0) - 1
    seg = torch.where(start[1:], torch.zeros_like(seg), seg)
    out = torch.zeros(int(run[-1]) + 1, dtype=seg.dtype, device=seg.device).index_add(0, run[1:], seg)
    return out, grid[start][:, :-1]


def weak_operator(op: torch.Tensor, points: torch.Tensor, weak_form) -> torch.Tensor:
    """`Operator._weak_pde_compute` (tedeous/eval.py:195-221): every residual column times the test functions,
    integrated over every grid column in turn.  op [N, n_eq] on the interior points -> [1, n_eq]."""
    cols = []
    for i in range(op.shape[-1]):
        sol = op[:, i]
        for func in weak_form:
            sol = sol * func(points).to(op.device).reshape(-1)
        g = points
        for _ in range(points.shape[-1]):
            sol, g = integration(sol, g)
        cols.append(sol.reshape(-1, 1))
    return cols[0] if len(cols) == 1 else torch.cat(cols).reshape(1, -1)


class Losses:
    def __init__(self, mode, weak_form, n_t, tol, n_t_operation=None):
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

    def _weak_loss(self, operator, bval, true_bval, lambda_op, lambda_bound):
        """losses.py:184-228: `operator` is already the integrated weak residual [1, n_eq]."""
        loss_bnd, diff = self._loss_bcs(bval, true_bval, lambda_bound)
        loss = operator @ lambda_op.T.to(operator) + loss_bnd
        with torch.no_grad():
            loss_normalized = operator.sum(1, keepdim=True) + diff.sum()
        return loss, loss_normalized

    def compute(self, operator, bval, true_bval, lambda_op, lambda_bound,
                save_graph: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
        if bval is None:
            raise ValueError('a problem without boundary conditions has no finite loss '
                             '(the reference prints a warning and returns inf, losses.py:250-255)')
        lambda_op = lambda_prepare(operator, lambda_op)
        lambda_bound = lambda_prepare(bval, lambda_bound)
        if self.weak_form not in (None, []):
            return self._weak_loss(operator, bval, true_bval, lambda_op, lambda_bound)
        if self.tol != 0:
            return self._causal_loss(operator, bval, true_bval, lambda_op, lambda_bound)
        return self._default_loss(operator, bval, true_bval, lambda_op, lambda_bound, save_graph)
