"""Mini-batch fixture from the UNMODIFIED reference (tedeous/eval.py:124-141, 174-182; solution.py:159-166):
wave equation, mode 'autograd', 21 x 21 grid, batch_size = 100 -> 5 batches per epoch (the last one has 41 rows).
Seven consecutive `evaluate(); backward()` calls (they cross the epoch boundary, where the rows are reshuffled) with the
DataLoader generator the reference creates itself (a fresh CPU torch.Generator: fixed default seed).
Run in the build container:  python tests/golden/make_minibatch.py"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, '/root/reference')
for name in ['SALib', 'matplotlib', 'matplotlib.pyplot', 'matplotlib.cm', 'seaborn']:
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules['SALib'].ProblemSpec = object

import problems  # noqa: E402
from tedeous import data as ref_data  # noqa: E402
from tedeous.device import solver_device  # noqa: E402
from tedeous.input_preprocessing import Operator_bcond_preproc  # noqa: E402
from tedeous.solution import Solution  # noqa: E402

solver_device('cpu')
prob = problems.wave(ref_data, 'float64', n=20, mode='autograd', layers=(2, 32, 32, 1))
grid = prob.domain.build('autograd')
net = problems.make_net(prob.net_layers, torch.float64, prob.init)
params = list(net.parameters())
weights = torch.cat([p.detach().reshape(-1) for p in params]).numpy()
bconds = prob.conditions.build(prob.domain.variable_dict)
eq = Operator_bcond_preproc(grid, prob.equation.equation_lst, bconds).set_strategy('autograd')
kw = prob.compile_kwargs
sol = Solution(grid, eq, net, 'autograd', None, kw['lambda_operator'], kw['lambda_bound'], batch_size=100)
losses, grads, rows = [], [], []
for _ in range(7):
    for p in params:
        p.grad = None
    batch = sol.operator.grid_batch.clone()
    loss, _ = sol.evaluate()
    loss.backward()
    losses.append(float(loss))
    grads.append(torch.cat([p.grad.reshape(-1) for p in params]).numpy())
    rows.append(batch.shape[0])
np.savez_compressed(os.path.join(HERE, 'minibatch_wave.npz'), weights=weights, losses=np.array(losses),
                    grads=np.stack(grads), rows=np.array(rows), n_batches=np.array(sol.operator.n_batches))
print('losses', losses, 'rows', rows, 'n_batches', sol.operator.n_batches)
