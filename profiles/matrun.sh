python -m pytest tests -m gpu -q -k mat 2>&1 | tail -4 > gpurun_out/s13_tests.log; cat gpurun_out/s13_tests.log
python bench.py --workload poisson_mat_4096 --no-cpu-baseline --steps 50 --warmup 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'], 'kernel_ms', d['roofline'].get('kernel_ms'), d['config']['kernel'])"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -k regex:'mat_' -s 12 -c 6 --csv --log-file gpurun_out/s13_launches_mat.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --workload poisson_mat_4096 > gpurun_out/s13_launches_mat.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/s13_launches_mat.csv')) if len(r)>10]
h=rows[0]; ik=h.index('Kernel Name'); im=h.index('Metric Name'); iv=h.index('Metric Value'); ii=h.index('ID')
for r in rows[1:]:
    print(r[ii], r[ik][:40], r[im], r[iv])
PY
