# quick check of the streamed tensor-core path: parity tests of the path, then the three workloads it carries
tag=${1:-q}
timeout 500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "streamed or auto_dispatch or large_grid_consistency or loss_and_gradient_match or weak_form or vector_jacobian or repeatable" 2>&1 | tail -3
for w in wave_autograd_1e6 ns_autograd_1e6 burgers_NN_1e6; do
  timeout 300 python bench.py --workload $w --no-cpu-baseline --steps 20 --warmup 5 2>gpurun_out/${tag}_$w.err > gpurun_out/${tag}_$w.json
  python -c "
import json
d=json.loads(open('gpurun_out/${tag}_$w.json').read().strip().splitlines()[-1]); print('$w ms/step %.4f value %.4g frac %.3f e2e %.4g launches %d' % (d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value'], d['gpu_launches']))" || tail -3 gpurun_out/${tag}_$w.err
done
