#!/bin/bash
# Run on the GPU box (gpurun): launch lists + full captures of the hot kernels for round $1.
R=${1:-r02}
mkdir -p gpurun_out
K="regex:tdb|jet_|wgrad|mat_|pack_|reduce_|fused_optimizer|optimizer_tick|peer_"
# every launch of OUR kernels during the default bench command (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv \
    --log-file gpurun_out/${R}_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/${R}_launches_bench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 200 --csv \
    --log-file gpurun_out/${R}_launches_mat.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --workload poisson_mat_4096 > gpurun_out/${R}_launches_mat.log 2>&1
# dominant kernels, --set full.  BASELINE config 1 (the bench headline) and the 10^6-point workloads all run on the
# streamed pair; -s skips the warm-up launches of the kernel
ncu --set full --clock-control none --import-source on -k regex:jet_tcs -s 3 -c 1 -o gpurun_out/${R}_jet_tcs_cfg1 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-configs --workload burgers_NN_cfg1 > gpurun_out/${R}_ncu_tcs_cfg1.log 2>&1
TDB200_NO_TC_BOUNDARY=1 ncu --set full --clock-control none --import-source on -k regex:jet_tcs -s 2 -c 1 -o gpurun_out/${R}_jet_tcs \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --workload wave_autograd_1e6 > gpurun_out/${R}_ncu_tcs.log 2>&1
TDB200_NO_TC_BOUNDARY=1 ncu --set full --clock-control none --import-source on -k regex:wgrad -s 2 -c 1 -o gpurun_out/${R}_wgrad \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --workload wave_autograd_1e6 > gpurun_out/${R}_ncu_wgrad.log 2>&1
TDB200_NO_TC_BOUNDARY=1 ncu --set full --clock-control none --import-source on -k regex:jet_tcs -s 2 -c 1 -o gpurun_out/${R}_jet_tcs_ns \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --workload ns_autograd_1e6 > gpurun_out/${R}_ncu_tcs_ns.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mat_march_kernel -s 2 -c 1 -o gpurun_out/${R}_mat \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --workload poisson_mat_4096 > gpurun_out/${R}_ncu_mat.log 2>&1
ls -la gpurun_out | tail -12
