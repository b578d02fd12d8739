for rep in 1 2 3; do for nb in 1 0; do
  if [ $nb = 1 ]; then export TDB200_NO_TC_BOUNDARY=1; else unset TDB200_NO_TC_BOUNDARY; fi
  timeout 200 python bench.py --workload wave_autograd_1e6 --no-cpu-baseline --steps 30 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('rep $rep no_tc_boundary=$nb ms/step %.4f clocks %s' % (d['ms_per_step'], d['clocks']))"
done; done
