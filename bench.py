#!/usr/bin/env python
"""Benchmark of the residual-loss hot path: collocation points per second for one loss + gradient step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

* A *step* is one evaluation of loss and the full parameter gradient over every collocation point
  (`Solution.evaluate(); loss.backward()` in the reference, one `tdb200_loss_grad` call here), plus the NCCL
  all-reduce of the [loss terms | gradient] vector when N > 1.
* Default workload = BASELINE.json configs[1]: wave equation 1D+t, mode 'autograd' (2nd-order jets), 10^6
  collocation points per GPU, tanh MLP 2-100-100-100-1.  Scaling is weak: every rank holds 10^6 points.
* `value` is timed with inputs resident in HBM; `e2e` goes through the public API (`Solution.evaluate` +
  `loss.backward()`) with the step's point set copied from pinned host memory and the result read back.
* `--impl reference` times the CPU port of the reference's algorithm (oracle/tedeous_oracle.py, torch CPU,
  all host threads) on a bounded sample of the same workload.

One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

METRIC = 'collocation pts/s for loss+grad step'
UNIT = 'pts/s'

# algorithmic FLOPs per point: 3 (fwd, bwd-data, bwd-weight) * J jet channels * 2 * sum(in*out)  
WORKLOADS = {
    # name: (builder name in tests/problems.py, kwargs, J, description)
    'wave_autograd_1e6': dict(fn='wave', kw=dict(n=999, mode='autograd', layers=(2, 100, 100, 100, 1)), J=5,
                              desc='wave 1D+t autograd, 1000x1000 pts/GPU, MLP 2-100-100-100-1'),
    'burgers_NN_cfg1': dict(fn='burgers', kw=dict(n=100, mode='NN', layers=(2, 100, 100, 100, 1)), J=4,
                            desc='Burgers 1D NN mode, 101x101 grid (9801 central pts), MLP 2-100-100-100-1'),
    'burgers_autograd_1e6': dict(fn='burgers', kw=dict(n=999, mode='autograd', layers=(2, 100, 100, 100, 1)), J=4,
                                 desc='Burgers 1D autograd, 1000x1000 pts/GPU'),
    'kdv_autograd_1e6': dict(fn='kdv', kw=dict(nx=999, nt=999, mode='autograd', layers=(2, 100, 100, 100, 1)), J=5,
                             desc='KdV periodic autograd, 1000x1000 pts/GPU'),
    'ns_autograd_1e6': dict(fn='navier_stokes', kw=dict(n=99, layers=(3, 100, 100, 100, 100, 100, 100, 3)), J=6,
                            desc='Navier-Stokes 2D+t autograd, 100^3 pts/GPU, MLP 3-100x6-3'),
    'poisson_mat_4096': dict(fn='poisson_mat', kw=dict(n=4095, derivative_points=2), J=0, mat=True,
                             desc="Poisson 2D mode 'mat', 4096x4096 grid, u_xx + u_yy - f, Dirichlet edges"),
}
MAT_BYTES_PER_CELL = 12      # read u, read the forcing tensor, write d loss / d u (fp32) - SURVEY 8d
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed `ncu --set full`
# captures (profiles/r01_ncu_jet_tc.md, profiles/r01_ncu_mat.md); None where no capture exists
NCU_TRAFFIC_BYTES = {'wave_autograd_1e6': 8.57e6 + 5.14e6, 'poisson_mat_4096': 147.86e6 + 35.75e6}


def flop_per_point(layers, J):
    return 3 * J * 2 * sum(a * b for a, b in zip(layers[:-1], layers[1:]))


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile('w', suffix='.csv', delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-i', str(self.idx), '-lms', '50'], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            lines = list(open(self.path))
            if not lines:                   # region shorter than the sampler's start-up: one query right after the load
                lines = subprocess.run(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-i',
                                        str(self.idx)], capture_output=True, text=True, timeout=10).stdout.splitlines()
            for line in lines:
                c = [x.strip() for x in line.split(',')]
                if len(c) < 9:
                    continue
                try:
                    sm.append(float(c[1])); mx.append(float(c[2]))
                except ValueError:
                    continue
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), c[5:9]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def load_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), 'measured'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback'


def measure_tf32_tflops(device, seconds=1.0):
    """Dense TF32 cuBLAS throughput of this GPU (the 3xTF32 roofline denominator is this / 3)."""
    n = 8192
    a = torch.randn(n, n, device=device)
    b = torch.randn(n, n, device=device)
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        for _ in range(3):
            a @ b
        torch.cuda.synchronize(device)
        best = 0.0
        t_end = time.time() + seconds
        while time.time() < t_end:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); a @ b; e1.record(); e1.synchronize()
            best = max(best, 2 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    del a, b
    return best


def make_problem(workload, api, world):
    """The workload with `world` times more nodes along the first axis (weak scaling: every rank keeps the
    single-GPU number of points)."""
    import problems
    spec = WORKLOADS[workload]
    kw = dict(spec['kw'])
    if world > 1:
        if 'nx' in kw:
            kw['nx'] = (kw['nx'] + 1) * world - 1
        else:
            kw['n0'] = (kw['n'] + 1) * world - 1
    return spec, getattr(problems, spec['fn'])(api, 'float32', **kw)


# ----------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch.distributed as dist
    import torch_de_solver_b200 as tdb
    import problems

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if args.gpus > 1 and world != args.gpus:
        raise SystemExit(f'--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    torch.set_default_device(dev)

    if WORKLOADS[args.workload].get('mat'):
        return run_b200_mat(args, dev, tdb, problems, rank, world)
    spec, prob = make_problem(args.workload, tdb, world)
    net = problems.make_net(prob.net_layers, torch.float32, prob.init).to(dev)
    model = tdb.Model(net, prob.domain, prob.equation, prob.conditions)
    model.compile(prob.mode, **prob.compile_kwargs, shard=(rank, world) if world > 1 else None)
    sol = model.solution_cls
    plan = sol._plan
    n_local = sol._ir.segments[0].n_groups
    n_global = sol._ir.n_interior
    params = list(net.parameters())

    def step_resident():
        out = plan.loss_grad()
        if world > 1:
            dist.all_reduce(out)
        return out

    graphed = False
    if world == 1 and n_local < 100_000 and not args.no_graph:      # launch-bound sizes: one CUDA graph per step
        try:
            step_resident, _ = plan.capture()
            graphed = True
        except Exception as e:                # noqa: BLE001 - report and time eager launches
            sys.stderr.write(f'[bench] CUDA graph capture failed ({e}); timing eager launches\n')

    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > 126 MB L2

    def timed(fn, steps):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for e0, e1 in ev:
            flush.zero_()
            e0.record()
            fn()
            e1.record()
        torch.cuda.synchronize(dev)
        return [e0.elapsed_time(e1) for e0, e1 in ev]

    sampler = ClockSampler(local)           # started before the warm-up: nvidia-smi needs ~0.2 s to deliver samples
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        step_resident()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t_wall = time.time()
    times = timed(step_resident, args.steps)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    clocks = sampler.stop()
    total_ms = sum(times)
    if world > 1:
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t)
    ms_per_step = total_ms / args.steps
    value = n_global / (ms_per_step * 1e-3)

    # ---- end to end through the public API, host buffers ----------------------------------------------
    flat = plan.flat
    host_in = [t.detach().cpu().pin_memory() for t in (flat.points, flat.targets, flat.coeffs)]
    dev_in = [flat.points, flat.targets, flat.coeffs]
    host_out = torch.empty(plan.out_size, dtype=torch.float32, device='cpu').pin_memory()
    h2d = sum(t.numel() * 4 for t in host_in)
    d2h = host_out.numel() * 4

    def step_e2e():
        for h, d in zip(host_in, dev_in):
            d.copy_(h, non_blocking=True)
        for p in params:
            p.grad = None
        loss, _ = sol.evaluate()
        loss.backward()
        host_out.copy_(sol._last_out, non_blocking=True)

    for _ in range(3):
        step_e2e()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e2e_steps = max(3, min(args.steps, 20))
    e2e_times = timed(step_e2e, e2e_steps)
    e2e_ms = sum(e2e_times)
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t)
    e2e_value = n_global / (e2e_ms / e2e_steps * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_src = load_peaks()
    torch.set_default_device('cpu')
    tf32 = measure_tf32_tflops(dev)
    fpp = flop_per_point(prob.net_layers, spec['J'])
    achieved = (n_local / (statistics.mean(times) * 1e-3)) * fpp / 1e12
    peak = tf32 / 3.0
    cpu = cpu_baseline(args.workload) if not args.no_cpu_baseline else None
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32' if plan.launches_per_call == 3 else 'tf32x3 (fp32 accumulate)', 'data': 'synthetic',
        'config': {'workload': args.workload, 'description': spec['desc'], 'points_per_gpu': n_local,
                   'points_total': n_global, 'mlp': list(prob.net_layers), 'mode': prob.mode,
                   'jet_channels': spec['J'], 'kernel': 'simt-fp32' if plan.launches_per_call == 3 else 'tcgen05-3xtf32 (interior) + simt-fp32 (boundary rows)',
                   'cuda_graph': graphed,
                   'l2': 'flushed between timed steps (256 MB write)', 'parallelism': f'dp{world} (points sharded)'},
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                'ms_per_step': e2e_ms / e2e_steps},
        'gpu_launches': args.steps * plan.launches_per_call,
        'clocks': clocks,
        'roofline': {'bound': 'tensor', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s',
                     'frac': achieved / peak if peak else None,
                     'traffic': NCU_TRAFFIC_BYTES.get(args.workload) if world == 1 else None,
                     'flop_per_point': fpp,
                     'peak_source': f'cuBLAS TF32 8192^3 measured live = {tf32:.1f} TFLOP/s, / 3 for 3xTF32 '
                                    f'(MEASURED_PEAKS.json [{peak_src}] bf16 = {peaks.get("bf16_tflops")})'},
        'cpu_baseline': cpu,
        'wall_s': time.time() - t_wall,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_b200_mat(args, dev, tdb, problems, rank=0, world=1):
    """mat-mode workload: HBM-bound stencil kernel.  Several GPUs: slab decomposition along axis 0, weak scaling
    (4096 rows per rank), 4 halo rows of u from each neighbour + one all-reduce of the loss terms per step."""
    import torch.distributed as dist
    from torch_de_solver_b200.mat import slab_rows
    spec = WORKLOADS[args.workload]
    kw = dict(spec['kw'])
    n1 = kw['n'] + 1
    if world > 1:
        kw['ny'] = kw['n']
        kw['n'] = (kw['n'] + 1) * world - 1
    prob = getattr(problems, spec['fn'])(tdb, 'float32', **kw)
    n0 = kw['n'] + 1
    r0, r1 = slab_rows(n0, rank, world)
    u = problems.make_mat_model((prob.mat_shape[0], r1 - r0, n1), torch.float32, seed=rank).to(dev).contiguous()
    model = tdb.Model(u, prob.domain, prob.equation, prob.conditions)
    model.compile('mat', **prob.compile_kwargs, shard=(rank, world) if world > 1 else None)
    sol = model.solution_cls
    plan = sol._plan
    n_cells = plan.n_cells                         # global
    n_local = plan.n_cells_local
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    def timed(fn, steps):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for e0, e1 in ev:
            flush.zero_()
            e0.record(); fn(); e1.record()
        torch.cuda.synchronize(dev)
        return [e0.elapsed_time(e1) for e0, e1 in ev]

    def reduce_max(ms_total):
        if world > 1:
            t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t)
        return ms_total

    step = lambda: plan.loss_grad_raw(u)
    graphed = False
    if not args.no_graph and world == 1:      # one CUDA graph per step (single rank; eager with collectives)
        try:
            step, _, _ = plan.capture(u)
            graphed = True
        except Exception as e:                # noqa: BLE001 - report and fall back to eager launches
            sys.stderr.write(f'[bench] CUDA graph capture failed ({e}); timing eager launches\n')
            step = lambda: plan.loss_grad_raw(u)
    sampler = ClockSampler(dev.index or 0)  # started before the warm-up: nvidia-smi needs ~0.2 s to deliver samples
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t_wall = time.time()
    times = timed(step, args.steps)
    if world > 1:
        dist.barrier()
    clocks = sampler.stop()
    ms = reduce_max(sum(times)) / args.steps
    # kernel-only time of the stencil kernel on this rank (roofline): the same launch without the exchange
    ue = u if world == 1 else torch.zeros(plan.ir.shape_ext, dtype=torch.float32, device=dev)
    # dominant (stencil) kernel alone: CUDA events recorded by the library on the launching stream around back-to-back
    # launches of that kernel (inputs + output = 192 MiB per launch > L2, so launches do not feed each other from cache)
    flush.zero_()
    kt = statistics.mean(plan.time_stencil(ue, 10) for _ in range(max(3, min(args.steps, 10))))
    # e2e: forcing tensor + boundary targets from pinned host memory every step, loss terms read back
    host_in = [plan._coeffs.detach().cpu().pin_memory(), plan._targets.detach().cpu().pin_memory()]
    dev_in = [plan._coeffs, plan._targets]
    host_out = torch.empty(plan.out_size, dtype=torch.float32, device='cpu').pin_memory()
    u.requires_grad_()

    def step_e2e():
        for h, d in zip(host_in, dev_in):
            d.copy_(h, non_blocking=True)
        u.grad = None
        loss, _ = sol.evaluate()
        loss.backward()
        host_out.copy_(sol._last_out, non_blocking=True)
    for _ in range(3):
        step_e2e()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e2e_steps = max(3, min(args.steps, 20))
    e2e_ms = reduce_max(sum(timed(step_e2e, e2e_steps))) / e2e_steps
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks, peak_src = load_peaks()
    achieved = n_local * MAT_BYTES_PER_CELL / (kt * 1e-3) / 1e9
    torch.set_default_device('cpu')
    cpu = cpu_baseline(args.workload) if (not args.no_cpu_baseline and world == 1) else None
    line = {
        'metric': METRIC, 'value': n_cells / (ms * 1e-3), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': args.workload, 'description': spec['desc'], 'cells': n_cells, 'cells_per_gpu': n_local,
                   'mode': 'mat', 'kernel': plan.kernel_kind, 'cuda_graph': graphed,
                   'l2': 'flushed between timed steps (256 MB write)',
                   'parallelism': 'single GPU' if world == 1 else
                                  f'{world} row slabs, {plan.ir.halo}-row halo exchange + all-reduce of the loss terms'},
        'e2e': {'value': n_cells / (e2e_ms * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': sum(t.numel() * 4 for t in host_in),
                'd2h_bytes_per_step': host_out.numel() * 4, 'ms_per_step': e2e_ms},
        'gpu_launches': args.steps * plan.launches_per_call,
        'clocks': clocks,
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                     'frac': achieved / peaks['hbm_gbs'],
                     'traffic': NCU_TRAFFIC_BYTES.get(args.workload) if world == 1 else None,
                     'bytes_per_cell': MAT_BYTES_PER_CELL, 'kernel_ms': kt, 'kernel': 'mat stencil kernel (' + plan.kernel_kind + ')',
                     'step_frac': n_local * MAT_BYTES_PER_CELL / (ms * 1e-3) / 1e9 / peaks['hbm_gbs'],
                     'peak_source': f'MEASURED_PEAKS.json [{peak_src}] hbm_gbs; achieved = cells per GPU * 12 B / mean '
                                    f'duration of the stencil kernel launch alone (CUDA events on its stream around 10 back-to-back '
                                    f'launches; working set 192 MiB > L2); step_frac = the same bytes / the whole step '
                                    f'(all launches, L2 flushed between steps)'},
        'cpu_baseline': cpu,
        'wall_s': time.time() - t_wall,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------
CPU_SAMPLE = {   # bounded CPU samples of the same operators (reference cost is flat in N)
    'wave_autograd_1e6': dict(fn='wave', kw=dict(n=315, mode='autograd', layers=(2, 100, 100, 100, 1))),
    'burgers_NN_cfg1': dict(fn='burgers', kw=dict(n=100, mode='NN', layers=(2, 100, 100, 100, 1))),
    'burgers_autograd_1e6': dict(fn='burgers', kw=dict(n=315, mode='autograd', layers=(2, 100, 100, 100, 1))),
    'kdv_autograd_1e6': dict(fn='kdv', kw=dict(nx=199, nt=199, mode='autograd', layers=(2, 100, 100, 100, 1))),
    'ns_autograd_1e6': dict(fn='navier_stokes', kw=dict(n=20, layers=(3, 100, 100, 100, 100, 100, 100, 3))),
    'poisson_mat_4096': dict(fn='poisson_mat', kw=dict(n=255, derivative_points=2)),
}


def cpu_step_time(workload, steps=3, warmup=1):
    """Times the oracle port (same call pattern as the reference) on the host cores."""
    import problems
    import torch_de_solver_b200 as tdb
    from oracle import tedeous_oracle as orc
    torch.set_default_device('cpu')
    spec = CPU_SAMPLE[workload]
    prob = getattr(problems, spec['fn'])(tdb, 'float32', **spec['kw'])
    grid = prob.domain.build(prob.mode)
    bconds = prob.conditions.build(prob.domain.variable_dict)
    kw = prob.compile_kwargs
    if prob.mode == 'mat':
        net = problems.make_mat_model(prob.mat_shape, torch.float32)
        params = [net.requires_grad_()]
    else:
        net = problems.make_net(prob.net_layers, torch.float32, prob.init)
        params = list(net.parameters())
    sol = orc.OracleSolution(grid, prob.equation.equation_lst, bconds, net, prob.mode, kw['lambda_operator'],
                             kw['lambda_bound'], h=kw.get('h', 0.001),
                             derivative_points=kw.get('derivative_points', 2))
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        orc.loss_and_grad(sol, params)
        ts.append(time.perf_counter() - t0)
    n = sol.op.shape[0]
    return n, ts[warmup:], f"{spec['fn']} {spec['kw']} -> {n} operator points"


def cpu_baseline(workload):
    n, ts, sample = cpu_step_time(workload)
    return {'value': n / min(ts), 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port',
            'sample': sample + f'; min of {len(ts)} steps after 1 warm-up; torch {torch.__version__} CPU, '
                               f'os.cpu_count()={os.cpu_count()}'}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    n, ts, sample = cpu_step_time(args.workload, steps=max(1, min(args.steps, 5)), warmup=min(max(args.warmup, 1), 2))
    ms = statistics.mean(ts) * 1e3
    val = n / (ms * 1e-3)
    spec = WORKLOADS[args.workload]
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': len(ts), 'warmup': min(max(args.warmup, 1), 2), 'ms_per_step': ms, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': args.workload, 'description': spec['desc'], 'sample': sample},
        'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port',
                         'sample': sample},
        'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='wave_autograd_1e6', choices=sorted(WORKLOADS))
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='mat workload: time eager launches instead of a CUDA graph')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
