#!/usr/bin/env python
"""Benchmark of the residual-loss hot path: collocation points per second for one loss + gradient step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME] [--no-configs]

* A *step* is one evaluation of loss and the full parameter gradient over every collocation point
  (`Solution.evaluate(); loss.backward()` in the reference, one `tdb200_loss_grad` call here), plus the NCCL
  all-reduce of the [loss terms | gradient] vector when N > 1.
* The top-level line is BASELINE.json's metric configuration: **Burgers 1D, mode 'NN', 101 x 101 grid, tanh MLP
  2-100-100-100-1** (configs[0], `burgers_NN_cfg1`).  Scaling is weak: with N ranks the grid has 101 N x 101 nodes
  and every rank holds a 101 x 101 block of it.
* The other BASELINE configs are measured in the same run, each with its own `value` / `e2e` / `roofline` /
  `cpu_baseline`, under the extra key `configs` (wave 10^6, Burgers NN 10^6, KdV 1.25 10^6 per GPU = 10^7 over 8,
  Navier-Stokes 1.25 10^6 per GPU = 10^7 over 8, Poisson mat 4096^2 per GPU).  `--workload NAME` measures one
  workload only (top level), `--no-configs` skips the extra ones.
* `value` is timed with inputs resident in HBM; `e2e` goes through the public API (`Solution.evaluate` +
  `loss.backward()`) with the step's inputs copied from pinned host memory and the result read back.
* `--impl reference` times the reference's own CPU implementation (`oracle/_ref`: the unmodified TEDEouS package,
  installed by oracle/make_ref.py; falls back to the oracle port when it is absent) with every host thread.

One JSON line on stdout (rank 0).
"""
import argparse
import contextlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

METRIC = 'collocation pts/s for loss+grad step (Burgers 1D, NN mode)'
UNIT = 'pts/s'
HEADLINE = 'burgers_NN_cfg1'
MLP3 = (2, 100, 100, 100, 1)

# name: builder in tests/problems.py, kwargs of the single-GPU problem, jet channels J, description.
# `axis0` names the kwarg of the slowest (sharded) axis: with N ranks it grows to (n + 1) N - 1 intervals.
WORKLOADS = {
    'burgers_NN_cfg1': dict(fn='burgers', kw=dict(n=100, mode='NN', layers=MLP3), J=4, axis0='n0',
                            desc='Burgers 1D NN mode (h = 0.001), 101x101 grid per GPU (9801 central pts), MLP 2-100-100-100-1'),
    'wave_autograd_1e6': dict(fn='wave', kw=dict(n=999, mode='autograd', layers=MLP3), J=5, axis0='n0',
                              desc='wave 1D+t autograd, 1000x1000 pts per GPU, MLP 2-100-100-100-1'),
    'burgers_NN_1e6': dict(fn='burgers', kw=dict(n=999, mode='NN', layers=MLP3), J=4, axis0='n0',
                           desc='Burgers 1D NN mode (h = 0.001), 1000x1000 grid per GPU (998x998 central pts)'),
    'burgers_autograd_1e6': dict(fn='burgers', kw=dict(n=999, mode='autograd', layers=MLP3), J=4, axis0='n0',
                                 desc='Burgers 1D autograd, 1000x1000 pts per GPU'),
    'kdv_autograd_1e7over8': dict(fn='kdv', kw=dict(nx=1249, nt=999, mode='autograd', layers=MLP3), J=5, axis0='nx',
                                  desc='KdV periodic autograd, 1250x1000 pts per GPU (10^4 x 10^3 = 10^7 over 8 GPUs)'),
    'ns_autograd_1e7over8': dict(fn='navier_stokes', kw=dict(n=214, n0=26, layers=(3, 100, 100, 100, 100, 100, 100, 3)),
                                 J=6, axis0='n0',
                                 desc='Navier-Stokes 2D+t autograd, 27x215x215 pts per GPU (216x215x215 = 10^7 over 8 GPUs), MLP 3-100x6-3'),
    'ns_autograd_1e6': dict(fn='navier_stokes', kw=dict(n=99, layers=(3, 100, 100, 100, 100, 100, 100, 3)), J=6, axis0='n0',
                            desc='Navier-Stokes 2D+t autograd, 100^3 pts per GPU, MLP 3-100x6-3'),
    'poisson_mat_4096': dict(fn='poisson_mat', kw=dict(n=4095, derivative_points=2), J=0, mat=True,
                             desc="Poisson 2D mode 'mat', 4096x4096 grid per GPU, u_xx + u_yy - f, Dirichlet edges"),
}
EXTRA_CONFIGS = ['wave_autograd_1e6', 'burgers_NN_1e6', 'kdv_autograd_1e7over8', 'ns_autograd_1e7over8', 'poisson_mat_4096']
MAT_BYTES_PER_CELL = 12      # read u, read the forcing tensor, write d loss / d u (fp32) - SURVEY 8d

# bounded CPU samples of the same operators for the reference arm (the reference's cost per point is flat in N,
# BASELINE.md 2); config 1 is timed at its full size
CPU_SAMPLE = {
    'burgers_NN_cfg1': dict(fn='burgers', kw=dict(n=100, mode='NN', layers=MLP3)),
    'wave_autograd_1e6': dict(fn='wave', kw=dict(n=315, mode='autograd', layers=MLP3)),
    'burgers_NN_1e6': dict(fn='burgers', kw=dict(n=140, mode='NN', layers=MLP3)),
    'burgers_autograd_1e6': dict(fn='burgers', kw=dict(n=315, mode='autograd', layers=MLP3)),
    'kdv_autograd_1e7over8': dict(fn='kdv', kw=dict(nx=199, nt=199, mode='autograd', layers=MLP3)),
    'ns_autograd_1e7over8': dict(fn='navier_stokes', kw=dict(n=20, layers=(3, 100, 100, 100, 100, 100, 100, 3))),
    'ns_autograd_1e6': dict(fn='navier_stokes', kw=dict(n=20, layers=(3, 100, 100, 100, 100, 100, 100, 3))),
    'poisson_mat_4096': dict(fn='poisson_mat', kw=dict(n=255, derivative_points=2)),
}


def flop_per_point(layers, J):
    """Algorithmic FLOPs per point: 3 (fwd, bwd-data, bwd-weight) * J jet channels * 2 * sum(in*out)  (SURVEY 8d)."""
    return 3 * J * 2 * sum(a * b for a, b in zip(layers[:-1], layers[1:]))


def ncu_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the workload's dominant kernel, from the committed
    `ncu --set full` captures: profiles/ncu_traffic.json, written by profiles/summarize.py next to the summaries."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')) as f:
            rec = json.load(f).get(workload)
        return None if rec is None else float(rec['bytes'])
    except Exception:                       # noqa: BLE001 - no capture: the key stays null
        return None


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


def workload_kwargs(workload, world):
    """kwargs of the problem with `world` times more nodes along the first (sharded) axis: weak scaling, every rank
    keeps the single-GPU number of points."""
    spec = WORKLOADS[workload]
    kw = dict(spec['kw'])
    if world > 1 and not spec.get('mat'):
        base = kw.get(spec['axis0']) or kw['n']
        kw[spec['axis0']] = (base + 1) * world - 1
    return spec, kw


def common_config(workload, world):
    """The part of `config` both arms print (the driver compares it)."""
    spec = WORKLOADS[workload]
    return {'workload': workload, 'description': spec['desc'], 'scaling_rule': 'weak: the grid grows along axis 0 with the '
            'number of GPUs, every rank holds the single-GPU block', 'n_gpus': world}


class Ctx:
    """Per-process state shared by the measurements of one run."""

    def __init__(self, args):
        import torch.distributed as dist
        self.args = args
        self.dist = dist
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.rank = int(os.environ.get('RANK', '0'))
        self.local = int(os.environ.get('LOCAL_RANK', '0'))
        if args.gpus > 1 and self.world != args.gpus:
            raise SystemExit(f'--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={self.world})')
        torch.cuda.set_device(self.local)
        self.dev = torch.device('cuda', self.local)
        if self.world > 1:
            dist.init_process_group('nccl', device_id=self.dev)
        self.flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=self.dev)   # > 126 MB L2
        self.peaks, self.peak_src = load_peaks()
        self._tf32 = None

    def tf32(self):
        if self._tf32 is None:
            torch.set_default_device('cpu')
            self._tf32 = measure_tf32_tflops(self.dev)
        return self._tf32

    def timed(self, fn, steps):
        """CUDA-event time of every step on the launching (current) stream, L2 flushed between steps."""
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for e0, e1 in ev:
            self.flush.zero_()
            e0.record()
            fn()
            e1.record()
        torch.cuda.synchronize(self.dev)
        return [e0.elapsed_time(e1) for e0, e1 in ev]

    def barrier(self):
        torch.cuda.synchronize(self.dev)
        if self.world > 1:
            self.dist.barrier()

    def max_over_ranks(self, x):
        if self.world > 1:
            t = torch.tensor([x], device=self.dev, dtype=torch.float64)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            return float(t)
        return x


# ----------------------------------------------------------------------------------------------------
def measure_nn(ctx, workload, steps, warmup, with_cpu):
    """One NN / autograd workload -> dict (value, ms_per_step, e2e, roofline, gpu_launches, config, clocks...)."""
    import torch_de_solver_b200 as tdb
    import problems
    args, dev, world, rank = ctx.args, ctx.dev, ctx.world, ctx.rank
    torch.set_default_device(dev)
    spec, kw = workload_kwargs(workload, world)
    prob = getattr(problems, spec['fn'])(tdb, 'float32', **kw)
    net = problems.make_net(prob.net_layers, torch.float32, prob.init).to(dev)
    model = tdb.Model(net, prob.domain, prob.equation, prob.conditions)
    model.compile(prob.mode, **prob.compile_kwargs, shard=(rank, world) if world > 1 else None)
    sol = model.solution_cls
    plan = sol._plan
    n_local = sol._ir.segments[0].n_groups
    n_global = sol._ir.n_interior
    params = list(net.parameters())

    def step_resident():
        return sol._run_plan()[0]               # launches + (N > 1) the all-reduce of [loss terms | gradient]

    graphed = False
    if n_local < 100_000 and not args.no_graph:          # launch-bound sizes: one CUDA graph per step
        try:
            step_resident, sol._graph_out = sol.capture_step()
            graphed = True
        except Exception as e:                # noqa: BLE001 - report and time eager launches
            sys.stderr.write(f'[bench] {workload}: CUDA graph capture failed ({e}); timing eager launches\n')
            step_resident = lambda: sol._run_plan()[0]

    sampler = ClockSampler(ctx.local)         # started before the warm-up: nvidia-smi needs ~0.2 s to deliver samples
    sampler.start()
    for _ in range(max(warmup, 3)):
        step_resident()
    ctx.barrier()
    t_wall = time.time()
    times = ctx.timed(step_resident, steps)
    ctx.barrier()
    clocks = sampler.stop()
    ms_per_step = ctx.max_over_ranks(sum(times)) / steps
    value = n_global / (ms_per_step * 1e-3)

    # ---- end to end through the public API, host buffers ----------------------------------------------
    # every step: the step's inputs (points, targets, coefficient buffers) from pinned host memory, the call a user
    # makes, the result vector back to pinned host memory.  The call is `Solution.evaluate(); loss.backward()` - or,
    # where the step is launch bound and captured, the replay of `Solution.capture_step()` (what Model.train replays,
    # with the optimiser update behind it).
    flat = plan.flat
    # (points, targets and coefficient buffers are views of one device buffer, `flat.inputs`: one copy per step)
    host_in = [flat.inputs.detach().cpu().pin_memory()]
    dev_in = [flat.inputs]
    host_out = torch.empty(plan.out_size, dtype=torch.float32, device='cpu').pin_memory()
    h2d = sum(t.numel() * 4 for t in host_in)
    d2h = host_out.numel() * 4
    if graphed:
        replay, g_out = step_resident, sol._graph_out

        def step_e2e():
            for h, d in zip(host_in, dev_in):
                d.copy_(h, non_blocking=True)
            replay()
            host_out.copy_(g_out, non_blocking=True)
    else:
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
    ctx.barrier()
    e2e_steps = max(3, min(steps, 20))
    e2e_ms = ctx.max_over_ranks(sum(ctx.timed(step_e2e, e2e_steps))) / e2e_steps
    ctx.barrier()
    e2e_value = n_global / (e2e_ms * 1e-3)
    # the whole training step as Model.train replays it (loss + gradient + fused Adam update), for the record
    train_ms = None
    if graphed and world == 1:
        try:
            from torch_de_solver_b200.optimizers.fused import FusedOptimizer, TrainStep
            ts = TrainStep(sol, FusedOptimizer('Adam', sol._ir.net.param_tensors(), lr=1e-5))
            for _ in range(3):
                ts.step()
            train_ms = sum(ctx.timed(ts.step, e2e_steps)) / e2e_steps
        except Exception as e:                # noqa: BLE001
            sys.stderr.write(f'[bench] {workload}: training-step graph failed ({e})\n')
    kernel_path = plan.kernel_path
    launches = plan.launches_per_call
    tc = not kernel_path.startswith('simt')
    del sol, model, plan
    torch.set_default_device('cpu')
    if rank != 0:
        return None
    tf32 = ctx.tf32()
    fpp = flop_per_point(prob.net_layers, spec['J'])
    achieved = (n_local / (statistics.mean(times) * 1e-3)) * fpp / 1e12
    peak = tf32 / 3.0
    cfg = common_config(workload, world)
    cfg.update({'points_per_gpu': n_local, 'points_total': n_global, 'mlp': list(prob.net_layers), 'mode': prob.mode,
                'jet_channels': spec['J'],
                'kernel': kernel_path + (' (interior, identity boundary rows) + simt-fp32 (other boundary rows)' if tc else ''),
                'cuda_graph': graphed, 'l2': 'flushed between timed steps (256 MB write)',
                'parallelism': f'dp{world} (points sharded, one all-reduce of [loss terms | gradient] per step)'})
    return {
        'value': value, 'unit': UNIT, 'ms_per_step': ms_per_step, 'steps': steps, 'warmup': max(warmup, 3),
        'dtype': 'tf32x3 (fp32 accumulate)' if tc else 'f32', 'config': cfg,
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                'ms_per_step': e2e_ms, 'call': 'Solution.capture_step() replay' if graphed else
                'Solution.evaluate(); loss.backward()', 'train_step_ms': train_ms},
        'gpu_launches': steps * launches, 'clocks': clocks,
        'roofline': {'bound': 'tensor', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s',
                     'frac': achieved / peak if peak else None,
                     'traffic': ncu_traffic(workload) if world == 1 else None, 'flop_per_point': fpp,
                     'peak_source': f'cuBLAS TF32 8192^3 measured live = {tf32:.1f} TFLOP/s, / 3 for 3xTF32 '
                                    f'(MEASURED_PEAKS.json [{ctx.peak_src}] bf16 = {ctx.peaks.get("bf16_tflops")})'},
        'cpu_baseline': cpu_baseline(workload) if with_cpu else None,
        'wall_s': time.time() - t_wall,
    }


def measure_mat(ctx, workload, steps, warmup, with_cpu):
    """mat-mode workload: HBM-bound stencil kernel.  Several GPUs: slab decomposition along axis 0, weak scaling
    (4096 rows per rank), halo rows of u from each neighbour + one all-reduce of the loss terms per step."""
    import torch_de_solver_b200 as tdb
    import problems
    from torch_de_solver_b200.mat import slab_rows
    args, dev, world, rank = ctx.args, ctx.dev, ctx.world, ctx.rank
    torch.set_default_device(dev)
    spec = WORKLOADS[workload]
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

    step = lambda: plan.loss_grad_raw(u)
    graphed = False
    if not args.no_graph:                     # one CUDA graph per step (halo exchange and all-reduce included)
        try:
            step, _, _ = plan.capture(u)
            graphed = True
        except Exception as e:                # noqa: BLE001 - report and fall back to eager launches
            sys.stderr.write(f'[bench] {workload}: CUDA graph capture failed ({e}); timing eager launches\n')
            step = lambda: plan.loss_grad_raw(u)
    sampler = ClockSampler(ctx.local)
    sampler.start()
    for _ in range(max(warmup, 3)):
        step()
    ctx.barrier()
    t_wall = time.time()
    times = ctx.timed(step, steps)
    ctx.barrier()
    clocks = sampler.stop()
    ms = ctx.max_over_ranks(sum(times)) / steps
    # dominant (stencil) kernel alone: CUDA events recorded by the library on the launching stream around back-to-back
    # launches of that kernel (inputs + output = 192 MiB per launch > L2, so launches do not feed each other from cache)
    ue = u if world == 1 else torch.zeros(plan.ir.shape_ext, dtype=torch.float32, device=dev)
    ctx.flush.zero_()
    kt = statistics.mean(plan.time_stencil(ue, 10) for _ in range(max(3, min(steps, 10))))
    # e2e through Solution.evaluate + backward: the step's inputs - the grid values u (the quantity a mat-mode optimiser
    # updates) and the boundary targets - come from pinned host memory, the loss terms are read back.  The forcing tensor
    # is static problem data: the reference's own step reads it from a resident tensor (tedeous/derivative.py:319-320),
    # so it is uploaded once, not per step.
    host_in = [u.detach().cpu().pin_memory(), plan._targets.detach().cpu().pin_memory()]
    dev_in = [u, plan._targets]
    host_out = torch.empty(plan.out_size, dtype=torch.float32, device='cpu').pin_memory()
    u.requires_grad_()

    def step_e2e():
        with torch.no_grad():
            for h, d in zip(host_in, dev_in):
                d.copy_(h, non_blocking=True)
        u.grad = None
        loss, _ = sol.evaluate()
        loss.backward()
        host_out.copy_(sol._last_out, non_blocking=True)
    for _ in range(3):
        step_e2e()
    ctx.barrier()
    e2e_steps = max(3, min(steps, 20))
    e2e_ms = ctx.max_over_ranks(sum(ctx.timed(step_e2e, e2e_steps))) / e2e_steps
    ctx.barrier()
    launches, kind, halo = plan.launches_per_call, plan.kernel_kind, plan.ir.halo
    h2d = sum(t.numel() * 4 for t in host_in)
    d2h = host_out.numel() * 4
    del sol, model, plan, u, ue
    torch.set_default_device('cpu')
    if rank != 0:
        return None
    peaks = ctx.peaks
    achieved = n_local * MAT_BYTES_PER_CELL / (kt * 1e-3) / 1e9
    cfg = common_config(workload, world)
    cfg.update({'cells': n_cells, 'cells_per_gpu': n_local, 'mode': 'mat', 'kernel': kind, 'cuda_graph': graphed,
                'l2': 'flushed between timed steps (256 MB write)',
                'parallelism': 'single GPU' if world == 1 else
                               f'{world} row slabs, {halo}-row halo exchange + all-reduce of the loss terms'})
    return {
        'value': n_cells / (ms * 1e-3), 'unit': UNIT, 'ms_per_step': ms, 'steps': steps, 'warmup': max(warmup, 3),
        'dtype': 'f32', 'config': cfg,
        'e2e': {'value': n_cells / (e2e_ms * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                'ms_per_step': e2e_ms,
                'note': 'u and the boundary targets are uploaded every step; the forcing tensor is static problem data '
                        '(resident in the reference too, tedeous/derivative.py:319-320)'},
        'gpu_launches': steps * launches, 'clocks': clocks,
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                     'frac': achieved / peaks['hbm_gbs'],
                     'traffic': ncu_traffic(workload) if world == 1 else None,
                     'bytes_per_cell': MAT_BYTES_PER_CELL, 'kernel_ms': kt, 'kernel': 'mat stencil kernel (' + kind + ')',
                     'step_frac': n_local * MAT_BYTES_PER_CELL / (ms * 1e-3) / 1e9 / peaks['hbm_gbs'],
                     'peak_source': f'MEASURED_PEAKS.json [{ctx.peak_src}] hbm_gbs; achieved = cells per GPU * 12 B / mean '
                                    f'duration of the stencil kernel launch alone (CUDA events on its stream around 10 '
                                    f'back-to-back launches; working set 192 MiB > L2); step_frac = the same bytes / the '
                                    f'whole step (all launches, L2 flushed between steps)'},
        'cpu_baseline': cpu_baseline(workload) if with_cpu else None,
        'wall_s': time.time() - t_wall,
    }


def measure(ctx, workload, steps, warmup, with_cpu):
    fn = measure_mat if WORKLOADS[workload].get('mat') else measure_nn
    return fn(ctx, workload, steps, warmup, with_cpu)


def run_b200(args):
    ctx = Ctx(args)
    with_cpu = not args.no_cpu_baseline and ctx.world == 1
    head = measure(ctx, args.workload, args.steps, args.warmup, with_cpu)
    configs = {}
    if args.workload == HEADLINE and not args.no_configs:
        for name in EXTRA_CONFIGS:
            try:
                res = measure(ctx, name, max(5, min(args.steps, 20)), max(3, min(args.warmup, 5)), with_cpu)
            except Exception as e:            # noqa: BLE001 - one failing extra config must not lose the headline line
                # (several ranks: a deterministic failure - an unsupported size, an allocation - hits every rank at the same
                # place, so the ranks stay in step; the error is reported in the line instead of a number)
                sys.stderr.write(f'[bench] rank {ctx.rank}: extra config {name} failed: {type(e).__name__}: {e}\n')
                res = {'error': f'{type(e).__name__}: {e}'}
            gc_collect()
            if ctx.rank == 0:
                configs[name] = res
    if ctx.rank == 0:
        line = {'metric': METRIC, 'value': head['value'], 'unit': UNIT, 'n_gpus': ctx.world, 'steps': args.steps,
                'warmup': max(args.warmup, 3), 'ms_per_step': head['ms_per_step'], 'higher_is_better': True,
                'scaling': 'weak', 'vs_baseline': None, 'dtype': head['dtype'], 'data': 'synthetic',
                'config': head['config'], 'e2e': head['e2e'], 'gpu_launches': head['gpu_launches'],
                'clocks': head['clocks'], 'roofline': head['roofline'], 'cpu_baseline': head['cpu_baseline'],
                'wall_s': head['wall_s']}
        if configs:
            line['configs'] = configs
        print(json.dumps(line))
    if ctx.world > 1:
        ctx.dist.destroy_process_group()


def gc_collect():
    import gc
    gc.collect()
    torch.cuda.empty_cache()


# ----------------------------------------------------------------------------------------------------
def _reference_solution(prob, grid, bconds, net):
    """The reference's own objects (oracle/_ref, unmodified TEDEouS): what Model.compile builds, tedeous/model.py:96-113."""
    from tedeous.input_preprocessing import Operator_bcond_preproc
    from tedeous.solution import Solution
    kw = prob.compile_kwargs
    eq_cls = Operator_bcond_preproc(grid, prob.equation.equation_lst, bconds, h=kw.get('h', 0.001),
                                    inner_order='1', boundary_order='2').set_strategy(prob.mode)
    return Solution(grid, eq_cls, net, prob.mode, kw.get('weak_form'), kw['lambda_operator'], kw['lambda_bound'],
                    tol=kw.get('tol', 0), derivative_points=kw.get('derivative_points', 2))


def cpu_step_time(workload, steps=3, warmup=1):
    """Times `loss, _ = Solution.evaluate(); loss.backward()` (tedeous/solution.py:129-168, closure.py:49-64) on the
    host cores: the installed reference (oracle/_ref) when present, else the oracle port of the same call pattern.
    -> (points, [seconds per step], sample description, kind, threads)."""
    import problems
    from oracle import make_ref
    torch.set_default_device('cpu')
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)            # torchrun exports OMP_NUM_THREADS=1: use every host core regardless
    spec = CPU_SAMPLE[workload]
    kind = 'reference' if make_ref.available() else 'port'
    if kind == 'reference':
        make_ref.import_reference()
        from tedeous import data as api
        from tedeous.device import solver_device
        solver_device('cpu')
    else:
        import torch_de_solver_b200 as api
    prob = getattr(problems, spec['fn'])(api, 'float32', **spec['kw'])
    grid = prob.domain.build(prob.mode)
    bconds = prob.conditions.build(prob.domain.variable_dict)
    kw = prob.compile_kwargs
    if prob.mode == 'mat':
        net = problems.make_mat_model(prob.mat_shape, torch.float32)
        params = [net.requires_grad_()]
    else:
        net = problems.make_net(prob.net_layers, torch.float32, prob.init)
        params = list(net.parameters())
    if kind == 'reference':
        sol = _reference_solution(prob, grid, bconds, net)

        def one_step():
            for p in params:
                p.grad = None
            loss, _ = sol.evaluate()
            loss.backward()
    else:
        from oracle import tedeous_oracle as orc
        sol = orc.OracleSolution(grid, prob.equation.equation_lst, bconds, net, prob.mode, kw['lambda_operator'],
                                 kw['lambda_bound'], h=kw.get('h', 0.001),
                                 derivative_points=kw.get('derivative_points', 2))
        one_step = lambda: orc.loss_and_grad(sol, params)
    ts = []
    for _ in range(warmup + steps):
        t0 = time.perf_counter()
        one_step()
        ts.append(time.perf_counter() - t0)
    n = sol.op.shape[0]
    what = 'unmodified TEDEouS 0.4.11 (oracle/_ref)' if kind == 'reference' else 'oracle port (oracle/_ref absent)'
    sample = (f"{spec['fn']} {spec['kw']} -> {n} operator points; {what}; torch {torch.__version__} CPU, "
              f'{threads} threads (os.cpu_count()={os.cpu_count()})')
    return n, ts[warmup:], sample, kind, threads


def cpu_baseline(workload):
    try:
        with contextlib.redirect_stdout(sys.stderr):       # the reference prints ("Default cpu processor is used.")
            n, ts, sample, kind, threads = cpu_step_time(workload)
    except Exception as e:                    # noqa: BLE001
        return {'error': f'{type(e).__name__}: {e}'}
    return {'value': n / min(ts), 'unit': UNIT, 'cores': threads, 'kind': kind,
            'sample': sample + f'; min of {len(ts)} steps after 1 warm-up'}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    world = int(os.environ.get('WORLD_SIZE', str(args.gpus)))
    full = args.workload == HEADLINE           # the headline config is small enough for the reference at full size
    steps = args.steps if full else max(1, min(args.steps, 5))
    warmup = args.warmup if full else min(max(args.warmup, 1), 2)
    with contextlib.redirect_stdout(sys.stderr):           # stdout carries exactly one JSON line
        n, ts, sample, kind, threads = cpu_step_time(args.workload, steps=steps, warmup=warmup)
    ms = statistics.mean(ts) * 1e3
    val = n / (ms * 1e-3)
    cfg = common_config(args.workload, world)
    cfg['sample'] = sample
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': len(ts), 'warmup': warmup, 'ms_per_step': ms, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': cfg,
        'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': threads, 'kind': kind, 'sample': sample},
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
    ap.add_argument('--workload', default=HEADLINE, choices=sorted(WORKLOADS))
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-configs', action='store_true', help='headline workload only (skip the `configs` key)')
    ap.add_argument('--no-graph', action='store_true', help='time eager launches instead of a CUDA graph')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
