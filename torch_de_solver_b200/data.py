"""User-facing problem declaration: Domain / Conditions / Equation.

Same signatures and semantics as tedeous/data.py (Domain 33-107, Conditions 110-323, Equation 326-339);
written from scratch.  Nothing here is on the hot path - it only produces the point lists and term dicts
that `plan.lower_problem` turns into the kernel IR."""
from typing import Dict, List, Union
import numpy as np
import torch

from .device import check_device
from .input_preprocessing import EquationMixin
from .data_CSG import Circle, Rectangle, csg_boundary, csg_difference

_DTYPES = {'float32': torch.float32, 'float64': torch.float64, 'float16': torch.float16}


def tensor_dtype(dtype):
    return _DTYPES.get(dtype, dtype)


class Domain:
    """Tensor-product grid builder (tedeous/data.py:33-107)."""

    def __init__(self, type='uniform'):
        self.type = type
        self.variable_dict = {}

    def variable(self, variable_name: str, variable_set: Union[List, torch.Tensor],
                 n_points: Union[None, int], dtype: str = 'float32') -> None:
        dtype = tensor_dtype(dtype)
        if isinstance(variable_set, torch.Tensor):
            self.variable_dict[variable_name] = check_device(variable_set).to(dtype)
        elif self.type == 'uniform':
            start, end = variable_set
            # n_points intervals -> n_points + 1 nodes (tedeous/data.py:63-66)
            self.variable_dict[variable_name] = torch.linspace(start, end, n_points + 1, dtype=dtype)

    def build(self, mode: str, removed_domains: list = None) -> torch.Tensor:
        axes = [v.cpu() for v in self.variable_dict.values()]
        if mode in ('autograd', 'NN'):
            if len(axes) == 1:
                return check_device(axes[0].reshape(-1, 1))
            grid = check_device(torch.cartesian_prod(*axes))
            for dom in (removed_domains or []):
                kind = list(dom.keys())[0]
                if kind == 'rectangle':
                    shape = Rectangle(dom[kind]['coords_min'], dom[kind]['coords_max'])
                elif kind == 'circle':
                    shape = Circle(dom[kind]['center'], dom[kind]['radius'])
                else:
                    raise ValueError(f'unknown removed domain {kind!r}')
                grid = csg_difference(grid, shape).detach().clone()
            return grid
        # mat mode: [d, N0, N1, ...] with 'ij' indexing (tedeous/data.py:101-103)
        return check_device(torch.stack(torch.meshgrid(*axes, indexing='ij')))


class Conditions:
    """Boundary / initial / data conditions (tedeous/data.py:110-323)."""

    def __init__(self):
        self.conditions_lst = []

    def _add(self, bnd, bop, bval, var, kind):
        self.conditions_lst.append({'bnd': bnd, 'bop': bop, 'bval': bval, 'var': var, 'type': kind})

    def dirichlet(self, bnd, value, var: int = 0):
        self._add(bnd, None, value, var, 'dirichlet')

    def operator(self, bnd, operator: dict, value):
        # the reference's attempt to read 'var' from the operator always fails (dict_keys is not
        # subscriptable, tedeous/data.py:151-154) so the stored var is always 0
        self._add(bnd, EquationMixin.equation_unify(operator), value, 0, 'operator')

    def periodic(self, bnd, operator: dict = None, var: int = 0):
        value = torch.tensor([0.])
        if operator is None:
            self._add(bnd, None, value, var, 'periodic')
        else:
            self._add(bnd, EquationMixin.equation_unify(operator), value, 0, 'periodic')

    def robin(self, bnd, value, operator: Dict = None, var: int = 0):
        self._add(bnd, EquationMixin.equation_unify(operator), value, var, 'robin')

    def data(self, bnd, operator, value, var: int = 0):
        if operator is not None:
            operator = EquationMixin.equation_unify(operator)
        self._add(bnd, operator, value, var, 'data')

    def _bnd_grid(self, bnd, variable_dict: dict, dtype) -> torch.Tensor:
        """Boundary sub-grid from a tensor or a {name: scalar | [lo, hi] | Tensor} dict
        (tedeous/data.py:241-286)."""
        if isinstance(bnd, torch.Tensor):
            out = check_device(bnd).to(dtype)
        else:
            if list(bnd.keys())[0] == 'circle':
                full = torch.cartesian_prod(*[variable_dict[v] for v in variable_dict])
                shape = Circle(bnd['circle']['center'], bnd['circle']['radius'])
                return csg_boundary(full, shape)
            cols = []
            for name in variable_dict:
                spec = bnd[name]
                if isinstance(spec, torch.Tensor):
                    cols.append(check_device(spec).to(dtype))
                elif isinstance(spec, (float, int)):
                    cols.append(check_device(torch.tensor([spec])).to(dtype))
                elif isinstance(spec, list):
                    axis = variable_dict[name]
                    cols.append(check_device(axis[(axis >= spec[0]) & (axis <= spec[1])]).to(dtype))
                else:
                    raise TypeError(f'bad boundary spec for {name!r}: {type(spec)}')
            out = torch.cartesian_prod(*cols).to(dtype)
        return out.reshape(-1, 1) if out.dim() == 1 else out

    def build(self, variable_dict: dict):
        if not self.conditions_lst:
            return None
        dtype = variable_dict[list(variable_dict.keys())[0]].dtype
        for cond in self.conditions_lst:
            if cond['type'] == 'periodic':
                cond['bnd'] = [self._bnd_grid(b, variable_dict, dtype) for b in cond['bnd']]
            else:
                cond['bnd'] = self._bnd_grid(cond['bnd'], variable_dict, dtype)
            val = cond['bval']
            if isinstance(val, torch.Tensor):
                cond['bval'] = check_device(val).to(dtype)
            elif isinstance(val, (float, int)):
                cond['bval'] = check_device(torch.ones_like(cond['bnd'][:, 0]) * val).to(dtype)
            elif callable(val):
                cond['bval'] = check_device(val(cond['bnd'])).to(dtype)
        return self.conditions_lst


class Equation:
    """Container of equations in operator-dict form (tedeous/data.py:326-339)."""

    def __init__(self):
        self.equation_lst = []

    def add(self, eq: dict):
        self.equation_lst.append(eq)
