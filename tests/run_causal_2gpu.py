"""Manual multi-GPU check (not collected by pytest): causal loss with the collocation rows sharded over the ranks of a
torchrun launch must reproduce the single-rank loss and gradient.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/run_causal_2gpu.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import problems  # noqa: E402
import torch_de_solver_b200 as tdb  # noqa: E402


def run(shard, dev):
    prob = problems.burgers(tdb, 'float32', n=31, mode='autograd', layers=(2, 32, 32, 1), tol=2.0)   # 32 time slices
    net = problems.make_net(prob.net_layers, torch.float32, prob.init).to(dev)
    model = tdb.Model(net, prob.domain, prob.equation, prob.conditions)
    model.compile(prob.mode, **prob.compile_kwargs, shard=shard)
    loss, _ = model.solution_cls.evaluate()
    loss.backward()
    return float(loss), torch.cat([p.grad.reshape(-1) for p in net.parameters()]).double().cpu().numpy()


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    torch.set_default_device(dev)
    loss_s, grad_s = run((rank, world), dev)
    if rank == 0:
        loss_1, grad_1 = run(None, dev)
        rel_l = abs(loss_s - loss_1) / abs(loss_1)
        rel_g = np.linalg.norm(grad_s - grad_1) / np.linalg.norm(grad_1)
        print(f'causal loss over {world} ranks: {loss_s:.8f} vs single rank {loss_1:.8f} (rel {rel_l:.2e}); grad rel err {rel_g:.2e}')
        assert rel_l < 2e-6 and rel_g < 2e-5
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
