import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import problems, torch_de_solver_b200 as tdb
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
def log(*a): print(f'[r{rank} {time.time() % 1000:.1f}]', *a, flush=True)
torch.cuda.set_device(local); dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
torch.set_default_device(dev)
log('pg ready')
prob = problems.wave(tdb, 'float32', n=63, mode='autograd')
net = problems.make_net(prob.net_layers, torch.float32, prob.init).to(dev)
model = tdb.Model(net, prob.domain, prob.equation, prob.conditions)
log('compiling')
model.compile(prob.mode, **prob.compile_kwargs, shard=(rank, world), collective='library')
log('compiled, has_comm', model.solution_cls._plan.has_comm)
out = model.solution_cls._run_plan()[0]
torch.cuda.synchronize(); log('run 1 done', float(out[0]))
out = model.solution_cls._run_plan()[0]
torch.cuda.synchronize(); log('run 2 done', float(out[0]))
replay, g = model.solution_cls.capture_step()
log('captured')
replay(); torch.cuda.synchronize(); log('replayed', float(g[0]))
dist.barrier(); log('barrier ok')
dist.destroy_process_group()
