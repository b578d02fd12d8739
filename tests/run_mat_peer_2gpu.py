"""Manual multi-GPU check (not collected by pytest): sharded mat mode with the peer-memory exchange (csrc/peer.cu) - the
row slabs of a torchrun launch must reproduce the single-rank loss and the single-rank gradient rows, eagerly and as a
captured CUDA graph, and agree with the NCCL path.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/run_mat_peer_2gpu.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import problems  # noqa: E402
import torch_de_solver_b200 as tdb  # noqa: E402
from torch_de_solver_b200.mat import slab_rows  # noqa: E402


def build(full, rows, shard, dev):
    prob = problems.poisson_mat(tdb, 'float32', n=255, ny=383, derivative_points=2)
    u = full[:, rows[0]:rows[1]].clone().to(dev).contiguous()
    model = tdb.Model(u, prob.domain, prob.equation, prob.conditions)
    model.compile('mat', **prob.compile_kwargs, shard=shard)
    return model.solution_cls, u


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    torch.set_default_device(dev)
    prob = problems.poisson_mat(tdb, 'float32', n=255, ny=383, derivative_points=2)
    full = problems.make_mat_model(prob.mat_shape, torch.float32, seed=0)
    n0 = prob.mat_shape[1]
    r0, r1 = slab_rows(n0, rank, world)
    sol1, u1 = build(full, (0, n0), None, dev)
    out1, grad1 = sol1._plan.loss_grad_raw(u1)
    res = {}
    for mode in ('peer', 'nccl'):
        os.environ['TDB200_MAT_COLLECTIVE'] = mode
        sol, u = build(full, (r0, r1), (rank, world), dev)
        plan = sol._plan
        assert (plan._peer is not None) == (mode == 'peer'), (mode, plan._peer)
        out, grad = plan.loss_grad_raw(sol.model)
        out2, grad2 = plan.loss_grad_raw(sol.model)              # a second step: parity buffers, sequence numbers
        torch.cuda.synchronize()
        assert torch.equal(out, out2) and torch.equal(grad, grad2)
        replay, g_out, g_grad = plan.capture(sol.model)
        for _ in range(3):
            replay()
        torch.cuda.synchronize()
        assert not plan.peer_error()
        assert torch.equal(g_out, out) and torch.equal(g_grad, grad), mode
        rel_l = abs(float(out[0]) - float(out1[0])) / abs(float(out1[0]))
        rel_g = float((grad - grad1[:, r0:r1]).norm() / grad1[:, r0:r1].norm())
        print(f'rank {rank}/{world} {mode}: loss {float(out[0]):.6f} vs single rank {float(out1[0]):.6f} (rel {rel_l:.2e}); '
              f'gradient rows rel err {rel_g:.2e}', flush=True)
        assert rel_l < 2e-6 and rel_g < 1e-5
        res[mode] = (out.clone(), grad.clone())
        # timing of the captured step
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier(); torch.cuda.synchronize()
        ev0.record()
        for _ in range(200):
            replay()
        ev1.record(); torch.cuda.synchronize()
        print(f'rank {rank} {mode}: {ev0.elapsed_time(ev1) / 200 * 1e3:.1f} us per captured step (256 x 384 grid)', flush=True)
        del replay, g_out, g_grad, plan, sol
    assert torch.allclose(res['peer'][0], res['nccl'][0], rtol=1e-6) and torch.equal(res['peer'][1], res['nccl'][1])
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
