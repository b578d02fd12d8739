timeout 400 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for w in burgers_NN_cfg1 wave_autograd_1e6; do
for ov in 0 1; do
  if [ $ov = 0 ]; then export TDB200_NO_OVERLAP=1; else unset TDB200_NO_OVERLAP; fi
  timeout 200 python bench.py --workload $w --no-cpu-baseline --steps 30 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$w overlap=$ov ms/step %.4f value %.4g frac %.3f e2e %.4g' % (d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value']))"
done; done
