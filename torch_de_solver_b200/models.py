"""Model helpers that shape the hot path's inputs (tedeous/models.py:183-226)."""
from typing import Any

import torch


def mat_model(domain: Any, equation: Any, nn_model: torch.nn.Module = None) -> torch.Tensor:
    """mat-mode "model": a [n_eq, N0, N1, ...] tensor of solution values (tedeous/models.py:198-226)."""
    grid = domain.build('mat')
    shape = [len(equation.equation_lst)] + list(grid.shape)[1:]
    if nn_model is not None:
        nn_grid = torch.vstack([grid[i].reshape(-1) for i in range(grid.shape[0])]).T.float()
        return nn_model(nn_grid).detach().reshape(shape)
    return torch.ones(shape)


def parameter_registr(model: torch.nn.Module, parameters: dict) -> None:
    """Register trainable equation coefficients on the net (inverse problems, tedeous/models.py:183-195)."""
    for key, value in parameters.items():
        parameters[key] = torch.nn.Parameter(torch.tensor([value], requires_grad=True).float())
        model.register_parameter(key, parameters[key])
