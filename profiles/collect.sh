#!/bin/bash
# Run on the GPU box (gpurun): launch list + full captures of the three hot kernels for round $1.
R=${1:-r01}
mkdir -p gpurun_out
# every launch of OUR kernels during the default bench command (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'tdb|jet_|mat_|pack_|reduce_' -c 200 --csv \
    --log-file gpurun_out/${R}_launches_wave.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_launches_wave.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'tdb|jet_|mat_|pack_|reduce_' -c 200 --csv \
    --log-file gpurun_out/${R}_launches_mat.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --workload poisson_mat_4096 > gpurun_out/${R}_launches_mat.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:jet_tc -s 2 -c 1 -o gpurun_out/${R}_jet_tc \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_ncu_tc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:jet_simt -s 2 -c 1 -o gpurun_out/${R}_jet_simt \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload ns_autograd_1e6 > gpurun_out/${R}_ncu_simt.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mat_march_kernel -s 2 -c 1 -o gpurun_out/${R}_mat \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --workload poisson_mat_4096 > gpurun_out/${R}_ncu_mat.log 2>&1
ls -la gpurun_out
