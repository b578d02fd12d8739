# A/B of the tcs weight placement on the same box: default (weights in tensor memory) vs TDB200_TCS_W_SMEM=1
for v in tm smem; do
  if [ $v = smem ]; then export TDB200_TCS_W_SMEM=1; else unset TDB200_TCS_W_SMEM; fi
  for w in wave_autograd_1e6 ns_autograd_1e6 burgers_NN_1e6; do
    timeout 300 python bench.py --workload $w --no-cpu-baseline --steps 20 --warmup 5 2>gpurun_out/ab_$w.err > gpurun_out/ab_$w.json
    python -c "
import json
d=json.loads(open('gpurun_out/ab_$w.json').read().strip().splitlines()[-1]); print('$v $w ms/step %.4f value %.4g frac %.3f' % (d['ms_per_step'], d['value'], d['roofline']['frac']))" || tail -3 gpurun_out/ab_$w.err
  done
done
