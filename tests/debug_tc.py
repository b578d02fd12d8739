"""Ad-hoc GPU diagnostic: tcgen05 path vs SIMT path vs golden for every eligible case (prints, never asserts)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch
import problems
import torch_de_solver_b200 as tdb
from helpers import load_golden, set_weights

torch.set_default_device('cuda:0')
names = sys.argv[1:] or [k for k in sorted(problems.ZOO) if 'mat' not in k and k not in ('navier_stokes_autograd', 'burgers_autograd_4h')]
for name in names:
    g = load_golden(name, 'float64')
    outs = {}
    for impl in (1, 2):
        prob = problems.ZOO[name](tdb, 'float32')
        net = problems.make_net(prob.net_layers, torch.float32, prob.init)
        set_weights(list(net.parameters()), g['weights'])
        net = net.to('cuda:0')
        m = tdb.Model(net, prob.domain, prob.equation, prob.conditions)
        m.compile(prob.mode, **prob.compile_kwargs, impl=impl)
        out = m.solution_cls._plan.loss_grad().double().cpu().numpy()
        torch.cuda.synchronize()
        outs[impl] = out
    k = 2 + m.solution_cls._n_slots
    gn = np.linalg.norm(g['grad'])
    for impl in (1, 2):
        o = outs[impl]
        print(f'{name:26s} impl={impl} loss={o[0]:.8g} (gold {float(g["loss"]):.8g}, rel {abs(o[0]-g["loss"])/abs(g["loss"]):.2e}) '
              f'grad rel err {np.linalg.norm(o[k:]-g["grad"])/gn:.2e} slots={o[2:k]}')
    # per-layer gradient error of tc vs simt
    off = 0
    errs = []
    for a, b in zip(prob.net_layers[:-1], prob.net_layers[1:]):
        for n in (a * b, b):
            d = outs[2][k + off:k + off + n] - outs[1][k + off:k + off + n]
            errs.append(np.linalg.norm(d) / (np.linalg.norm(outs[1][k + off:k + off + n]) + 1e-30))
            off += n
    print('   per-tensor rel diff tc vs simt:', ' '.join(f'{e:.1e}' for e in errs))
