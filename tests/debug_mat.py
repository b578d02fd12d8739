"""Ad-hoc GPU diagnostic: mat-mode register-tap kernel vs the generic kernel vs an fp64 dense evaluation."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch
import problems
import torch_de_solver_b200 as tdb
from torch_de_solver_b200.mat import derivative_band
from test_mat_cpu import dense_from_band

torch.set_default_device('cuda:0')
n = int(sys.argv[1]) if len(sys.argv) > 1 else 511
prob = problems.poisson_mat(tdb, 'float32', n=n)
x = torch.linspace(0, 1, n + 1)
u = (torch.sin(np.pi * x)[:, None] * torch.sin(np.pi * x)[None, :]).reshape(1, n + 1, n + 1).contiguous()
u0 = (u + 0.05 * torch.sin(5 * np.pi * x)[:, None] * torch.sin(np.pi * x)[None, :]).contiguous()
res = {}
for tag, env in (('lin1', None), ('generic', '1')):
    if env: os.environ['TDB200_MAT_NO_LIN1'] = env
    else: os.environ.pop('TDB200_MAT_NO_LIN1', None)
    model = tdb.Model(u.clone(), prob.domain, prob.equation, prob.conditions)
    model.compile('mat', **prob.compile_kwargs)
    out, grad = model.solution_cls._plan.loss_grad_raw(u0)
    torch.cuda.synchronize()
    res[tag] = (out.cpu().numpy().copy(), grad[0].double().cpu().numpy().copy())
    print(tag, 'out', res[tag][0])
h = float(x[1] - x[0])
band, b, E = derivative_band(n + 1, 2, 2, h)
D2 = torch.as_tensor(dense_from_band(band.astype(np.float64), b, E, n + 1), dtype=torch.float64)
f64 = (-2 * np.pi ** 2 * torch.sin(np.pi * x.double())[:, None] * torch.sin(np.pi * x.double())[None, :])
U = u0[0].double()
r = D2 @ U + U @ D2.T - f64
N = float((n + 1) ** 2)
tgt = torch.zeros_like(U); tgt[:, -1] = torch.sin(np.pi * x.double())
cnt = torch.zeros_like(U)
for sl in ((0, slice(None)), (-1, slice(None)), (slice(None), 0), (slice(None), -1)):
    cnt[sl] += 1
n_b = 4 * (n + 1)
gref = (2.0 / N * (D2.T @ r + r @ D2) + 100.0 * 2.0 / n_b * cnt * (U - tgt)).cpu().numpy()
for tag in res:
    g = res[tag][1]
    d = np.abs(g - gref)
    i = np.unravel_index(np.argmax(d), d.shape)
    print(tag, 'rel err', np.linalg.norm(g - gref) / np.linalg.norm(gref), 'max abs diff', d.max(), 'at', i, 'g', g[i], 'ref', gref[i])
    rows = np.where(d.max(axis=1) > 1e-3 * np.abs(gref).max())[0]
    cols = np.where(d.max(axis=0) > 1e-3 * np.abs(gref).max())[0]
    print('   rows with large error:', rows[:20], '... n =', len(rows), ' cols:', cols[:20], '... n =', len(cols))
d = np.abs(res['lin1'][1] - res['generic'][1])
print('lin1 vs generic max abs diff', d.max(), 'at', np.unravel_index(np.argmax(d), d.shape))
