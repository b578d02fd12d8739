"""Manual multi-GPU check (not collected by pytest): the all-reduce under the C ABI (tdb200_plan_comm_init) - rows
sharded over the ranks of a torchrun launch must reproduce the single-rank loss and gradient, eagerly and as a captured
CUDA graph, and agree with the torch.distributed collective.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/run_comm_2gpu.py
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


def build(shard, dev, collective='library', name='wave'):
    prob = getattr(problems, name)(tdb, 'float32', n=63, mode='autograd')
    net = problems.make_net(prob.net_layers, torch.float32, prob.init).to(dev)
    model = tdb.Model(net, prob.domain, prob.equation, prob.conditions)
    model.compile(prob.mode, **prob.compile_kwargs, shard=shard, collective=collective)
    return model.solution_cls


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    torch.set_default_device(dev)
    sol_lib = build((rank, world), dev, 'library')
    assert sol_lib._plan.has_comm and sol_lib._plan._peer is None
    out_lib = sol_lib._run_plan()[0].double().cpu().numpy()
    os.environ['TDB200_COLLECTIVE'] = 'peer'              # the all-reduce over CUDA-IPC mapped peer memory (csrc/peer.cu)
    sol_peer = build((rank, world), dev, 'library')
    assert sol_peer._plan.has_comm and sol_peer._plan._peer is not None
    out_nccl = sol_peer._run_plan()[0].double().cpu().numpy()
    os.environ.pop('TDB200_COLLECTIVE')
    sol_t = build((rank, world), dev, 'torch')
    assert not sol_t._plan.has_comm
    out_t = sol_t._run_plan()[0].double().cpu().numpy()
    replay, g_out = sol_lib.capture_step()
    replay(); replay()
    torch.cuda.synchronize()
    out_graph = g_out.double().cpu().numpy()
    if rank == 0:
        out_1 = build(None, dev)._run_plan()[0].double().cpu().numpy()
        k = 2 + sol_lib._n_slots
        for what, o in (('library', out_lib), ('library (peer memory)', out_nccl), ('torch.distributed', out_t), ('library, CUDA graph', out_graph)):
            rel_l = abs(o[0] - out_1[0]) / abs(out_1[0])
            rel_g = np.linalg.norm(o[k:] - out_1[k:]) / np.linalg.norm(out_1[k:])
            print(f'{world} ranks, {what}: loss {o[0]:.8f} vs single rank {out_1[0]:.8f} (rel {rel_l:.2e}); grad rel err {rel_g:.2e}')
            assert rel_l < 2e-6 and rel_g < 2e-5
        assert np.array_equal(out_lib, out_graph)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
