python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for w in burgers_NN_cfg1 burgers_autograd_1e6 kdv_autograd_1e6 ns_autograd_1e6 wave_autograd_1e6 poisson_mat_4096; do
  timeout 300 python bench.py --workload $w --no-cpu-baseline --steps 20 --warmup 5 2>gpurun_out/s17_$w.err > gpurun_out/s17_$w.json
  python -c "
import json
d=json.loads(open('gpurun_out/s17_$w.json').read().strip().splitlines()[-1]); print('$w ms/step %.4f value %.4g frac %.3f e2e %.4g launches %d kernel %s' % (d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value'], d['gpu_launches'], d['config'].get('kernel')))" || tail -3 gpurun_out/s17_$w.err
done
