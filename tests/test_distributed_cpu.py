"""Multi-rank host logic on the CPU (gloo, world_size 2): every rank lowers its shard of the problem
(plan.lower_problem(shard=(rank, world))), evaluates it (tests/ir_interp.py stands in for the kernel), and the
ranks all-reduce the [loss | gradient] vector exactly as Solution._run_plan does on GPUs.  The reduced vector must
equal the single-rank result: slot denominators are global, so partial sums simply add."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import problems  # noqa: F401  (sys.path set by conftest)
from ir_interp import evaluate_ir
from test_lowering_cpu import lower


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, name, out_path):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        g, prob, model, ir = lower(name, shard=(rank, world))
        loss, loss_n, mse, _ = evaluate_ir(ir, model)
        grads = torch.autograd.grad(loss, list(model.parameters()))
        vec = torch.cat([loss.reshape(1), loss_n.reshape(1), torch.stack(mse)] + [x.reshape(-1) for x in grads]).detach()
        dist.all_reduce(vec, op=dist.ReduceOp.SUM)
        rows = sum(s.n_groups for s in ir.segments)
        cnt = torch.tensor([rows], dtype=torch.int64)
        dist.all_reduce(cnt)
        if rank == 0:
            np.savez(out_path, vec=vec.numpy(), rows=int(cnt))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('name', ['kdv_autograd', 'nonlinear_mix_NN'])
def test_two_ranks_reduce_to_single_rank_result(name, tmp_path):
    world = 2
    out = str(tmp_path / 'out.npz')
    mp.spawn(_worker, args=(world, _free_port(), name, out), nprocs=world, join=True)
    got = np.load(out)
    g, prob, model, ir = lower(name)
    loss, loss_n, mse, _ = evaluate_ir(ir, model)
    grads = torch.autograd.grad(loss, list(model.parameters()))
    ref = torch.cat([loss.reshape(1), loss_n.reshape(1), torch.stack(mse)] + [x.reshape(-1) for x in grads]).detach().numpy()
    assert int(got['rows']) == sum(s.n_groups for s in ir.segments)       # every row owned by exactly one rank
    np.testing.assert_allclose(got['vec'], ref, rtol=1e-11, atol=1e-13)
    assert float(got['vec'][0]) == pytest.approx(float(g['loss']), rel=1e-9 if prob.mode == 'autograd' else 1e-5)


def test_shard_ranges_partition_rows():
    for world in (2, 3, 8):
        seen = {}
        for rank in range(world):
            g, prob, model, ir = lower('wave_autograd', shard=(rank, world))
            for s in ir.segments:
                lo, hi = s.shard_range
                seen.setdefault(s.name, []).append((lo, hi, s.n_groups_global))
        for name, ranges in seen.items():
            ranges.sort()
            assert ranges[0][0] == 0 and ranges[-1][1] == ranges[0][2]
            for (a0, a1, _), (b0, b1, _) in zip(ranges[:-1], ranges[1:]):
                assert a1 == b0
