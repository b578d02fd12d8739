"""Debug script (not a test): impl=3 against impl=1 on a few problems, with per-block gradient errors."""
import sys
import numpy as np
import torch
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import problems
import torch_de_solver_b200 as tdb
from helpers import load_golden, set_weights

torch.set_default_device('cuda:0')
names = sys.argv[1:] or ['wave_autograd', 'kdv_autograd', 'burgers_autograd_4h', 'navier_stokes_autograd']
for name in names:
    g = load_golden(name, 'float64')
    outs = {}
    for impl in (1, 3):
        prob = problems.ZOO[name](tdb, 'float32')
        net = problems.make_net(prob.net_layers, torch.float32, prob.init)
        set_weights(list(net.parameters()), g['weights'])
        net = net.to('cuda:0')
        model = tdb.Model(net, prob.domain, prob.equation, prob.conditions)
        import os
        kw = dict(prob.compile_kwargs)
        if os.environ.get('LB'):
            kw['lambda_bound'] = float(os.environ['LB'])
        if os.environ.get('LO'):
            kw['lambda_operator'] = float(os.environ['LO'])
        model.compile(prob.mode, **kw, impl=impl)
        sol = model.solution_cls
        out = sol._run_plan()[0].double().cpu().numpy()
        torch.cuda.synchronize()
        outs[impl] = out
        k = 2 + sol._n_slots
    a, b = outs[1], outs[3]
    print(name, 'loss', a[0], b[0], 'rel', abs(a[0] - b[0]) / abs(a[0]), 'golden', float(g['loss']))
    ga, gb = a[k:], b[k:]
    print('  grad rel err', np.linalg.norm(ga - gb) / np.linalg.norm(ga), 'vs golden', np.linalg.norm(gb - g['grad']) / np.linalg.norm(g['grad']))
    off = 0
    layers = prob.net_layers
    for l, (i, o) in enumerate(zip(layers[:-1], layers[1:])):
        for nm, sz in (('W', i * o), ('b', o)):
            x, y = ga[off:off + sz], gb[off:off + sz]
            print(f'    {nm}{l}: |ref| {np.linalg.norm(x):.3e} err {np.linalg.norm(x - y):.3e}')
            off += sz
