"""Per-fixture parity table (not a test): fused CUDA path vs the fp64 goldens of the unmodified reference, for every
kernel choice.  Run on a GPU box:  python tests/parity_report.py > profiles/r02_parity.md"""
import sys

import numpy as np
import torch

sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import problems  # noqa: E402
import torch_de_solver_b200 as tdb  # noqa: E402
from helpers import load_golden, set_weights  # noqa: E402

torch.set_default_device('cuda:0')
IMPL = {0: 'auto', 1: 'SIMT fp32', 2: 'tcgen05 (dW in TMEM)', 3: 'tcgen05 streamed'}
print('# Parity of the fused CUDA path against the fp64 fixtures of the unmodified reference\n')
print('Tolerances of the north star: loss <= 1e-5 relative, gradient norm <= 1e-4 relative.  `grad vec` = |g - g_ref| / |g_ref|.')
print('NN-mode rows: the interior uses exact jets where the reference uses central differences of step h (DESIGN 5), so')
print('`op mse` carries the O(h^2) truncation of the REFERENCE (2e-4 at h = 0.01); loss and gradient stay inside the tolerance.\n')
print('| fixture | mode | points | kernel | loss rel err | grad norm rel err | grad vec rel err | op mse max rel err |')
print('|---|---|---|---|---|---|---|---|')
for name in sorted(problems.ZOO):
    g = load_golden(name, 'float64')
    prob0 = problems.ZOO[name](tdb, 'float32')
    impls = [0] if prob0.mode == 'mat' or prob0.compile_kwargs.get('weak_form') else [0, 1, 2, 3]
    for impl in impls:
        prob = problems.ZOO[name](tdb, 'float32')
        try:
            if prob.mode == 'mat':
                net = torch.as_tensor(g['weights']).reshape(prob.mat_shape).float().to('cuda:0').contiguous().requires_grad_()
                params = [net]
            else:
                net = problems.make_net(prob.net_layers, torch.float32, prob.init)
                set_weights(list(net.parameters()), g['weights'])
                net = net.to('cuda:0')
                params = list(net.parameters())
            model = tdb.Model(net, prob.domain, prob.equation, prob.conditions)
            opts = {'impl': impl} if impl else {}
            model.compile(prob.mode, **prob.compile_kwargs, **opts)
        except RuntimeError as e:
            if impl:
                continue          # this kernel does not serve this net (e.g. impl = 2 on deep nets)
            raise
        sol = model.solution_cls
        loss, _ = sol.evaluate()
        loss.backward()
        grad = torch.cat([p.grad.reshape(-1) for p in params]).double().cpu().numpy()
        gn = np.linalg.norm(g['grad'])
        el = abs(float(loss) - float(g['loss'])) / abs(float(g['loss']))
        en = abs(np.linalg.norm(grad) - gn) / gn
        ev = np.linalg.norm(grad - g['grad']) / gn
        try:
            om = sol.op_mse.double().cpu().numpy()
            eo = float(np.max(np.abs(om - g['op_mse']) / (np.abs(g['op_mse']) + 1e-300))) if om.shape == g['op_mse'].shape else float('nan')
        except Exception:           # noqa: BLE001
            eo = float('nan')
        kern = IMPL[impl] if prob.mode != 'mat' else sol._plan.kernel_kind
        print(f'| {name} | {prob.mode} | {int(g["op_rows"])} | {kern} | {el:.1e} | {en:.1e} | {ev:.1e} | {eo:.1e} |')
