"""Natural-gradient descent on the fused path (tedeous/optimizers/ngd.py, https://arxiv.org/abs/2302.13163).

The reference assembles the Gram matrix `1/len(r) J^T J` of the interior and the boundary residuals from one
`torch.autograd.grad` call PER RESIDUAL (ngd.py:57-77).  Here the per-residual Jacobian rows come from
`Solution.residual_jacobian()` (tdb200_jacobian_rows: one kernel launch per segment and residual column), the Gram
products and the pseudo-inverse solve are dense library calls on the device, and the 31 loss evaluations of the grid line
search (ngd.py:29-55) are forward launches of the fused plan.
"""
import torch
from torch.nn.utils import parameters_to_vector, vector_to_parameters


class NGD(torch.optim.Optimizer):
    def __init__(self, params, grid_steps_number: int = 30):
        super().__init__(params, {'grid_steps_number': grid_steps_number})
        self.params = self.param_groups[0]['params']
        self.grid_steps_number = grid_steps_number
        self.steps = 0.5 ** torch.linspace(0, grid_steps_number, grid_steps_number + 1)

    @staticmethod
    def gram(jac: torch.Tensor) -> torch.Tensor:
        """1 / len(residuals) * J^T J (ngd.py:57-77; zero rows - the padding of `bval` - count in the length)."""
        return jac.T @ jac / max(jac.shape[0], 1)

    @staticmethod
    def pinv_solve(A: torch.Tensor, b: torch.Tensor, tol: float = None) -> torch.Tensor:
        """Least-squares solution through the SVD with singular values <= tol dropped (ngd.py:104-126)."""
        tol = torch.finfo(A.dtype).eps if tol is None else tol
        U, S, Vh = torch.linalg.svd(A, full_matrices=False)
        Sinv = torch.where(S > tol, 1.0 / S, torch.zeros_like(S))
        return Vh.mT @ (Sinv * (U.mT @ b))

    def _line_search(self, solution, nat_grad: torch.Tensor) -> None:
        base = parameters_to_vector(self.params).detach().clone()
        losses = []
        with torch.no_grad():
            for step in self.steps.tolist():
                vector_to_parameters(base - step * nat_grad, self.params)
                loss, _ = solution.evaluate(save_graph=False)
                losses.append(loss.reshape(1).detach())
            best = self.steps[int(torch.argmin(torch.cat(losses)))].item()
            vector_to_parameters(base - best * nat_grad, self.params)

    def step(self, closure=None) -> torch.Tensor:
        """closure() -> (int_res, bval, true_bval, loss, loss_function) like `Closure._closure_ngd`
        (closure.py:98-116); `loss_function` is the bound `Solution.evaluate`, whose Solution supplies the Jacobians."""
        _, _, _, loss, loss_function = closure()
        solution = loss_function.__self__
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
        f_grads = parameters_to_vector(grads).detach()
        j_int, j_bnd = solution.residual_jacobian()
        G = self.gram(j_int) + self.gram(j_bnd)
        # Marquardt-Levenberg term of the reference: min(loss, 0) * Id (zero for a positive loss, ngd.py:176-178)
        damp = torch.clamp(loss.detach().reshape(()), max=0.0)
        if float(damp) != 0.0:
            G = G + damp * torch.eye(G.shape[0], device=G.device, dtype=G.dtype)
        nat_grad = self.pinv_solve(G, f_grads)
        self._line_search(solution, nat_grad)
        return loss
