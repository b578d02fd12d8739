for EB in 148 296 444 148 296 444; do
  TDB200_MARCH_EDGE_BLOCKS=$EB timeout 200 python bench.py --workload poisson_mat_4096 --no-cpu-baseline --steps 100 --warmup 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('EB=$EB ms/step %.4f kernel_ms %.4f' % (d['ms_per_step'], d['roofline']['kernel_ms']))"
done
