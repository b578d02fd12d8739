"""Debug script (not a test): where the multi-rank mat step spends its time (4096 x 4096 cells per rank)."""
import os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import problems
import torch_de_solver_b200 as tdb
from torch_de_solver_b200.mat import slab_rows

rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
torch.set_default_device(dev)
n1 = 4095
prob = problems.poisson_mat(tdb, 'float32', n=4096 * world - 1, ny=n1, derivative_points=2)
n0 = prob.mat_shape[1]
r0, r1 = slab_rows(n0, rank, world)
u = problems.make_mat_model((1, r1 - r0, n1 + 1), torch.float32, seed=rank).to(dev).contiguous()
model = tdb.Model(u, prob.domain, prob.equation, prob.conditions)
model.compile('mat', **prob.compile_kwargs, shard=(rank, world) if world > 1 else None)
plan = model.solution_cls._plan
print(rank, 'kind', plan.kernel_kind, 'launches', plan.launches_per_call, 'peer', plan._peer is not None, 'ext', plan.ir.shape_ext, flush=True)
t0 = time.time()
replay, out, grad = plan.capture(u)
torch.cuda.synchronize()
print(rank, 'capture took %.2f s' % (time.time() - t0), 'peer_error', plan.peer_error(), flush=True)
for _ in range(5):
    replay()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(100):
    replay()
e1.record(); torch.cuda.synchronize()
print(rank, 'captured step back to back: %.1f us' % (e0.elapsed_time(e1) * 10), 'peer_error', plan.peer_error(), flush=True)
ue = u if world == 1 else plan._ext
print(rank, 'stencil kernel alone: %.1f us' % (plan.time_stencil(ue, 20) * 1e3), flush=True)
# eager pieces
def timeit(f, n=50):
    for _ in range(3): f()
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
print(rank, 'loss_grad_ext eager: %.1f us' % timeit(lambda: plan.loss_grad_ext(ue)), flush=True)
if world > 1 and plan._peer is not None:
    import ctypes as C
    ir = plan.ir; n_var, n_ext, nn1 = ir.shape_ext; up, n, h = ir.rows[0] - ir.ext[0], ir.rows[1] - ir.rows[0], ir.halo
    st = torch.cuda.current_stream(dev).cuda_stream
    o = torch.zeros(plan.out_size, device=dev)
    dist.barrier()
    print(rank, 'halo kernel eager: %.1f us' % timeit(lambda: plan.lib.tdb200_peer_halo(plan._peer, plan._ext.data_ptr(), n_ext * nn1, n_var, h * nn1, up * nn1, (up + n - h) * nn1, 0, (up + n) * nn1, st)), flush=True)
    dist.barrier()
    print(rank, 'loss kernel eager: %.1f us' % timeit(lambda: plan.lib.tdb200_peer_allreduce(plan._peer, o.data_ptr(), plan.out_size, st)), 'peer_error', plan.peer_error(), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
