set -x
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python bench.py > gpurun_out/s14_bench_wave.json 2> gpurun_out/s14_bench_wave.err; tail -c 1500 gpurun_out/s14_bench_wave.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s14_bench_ref.json 2> gpurun_out/s14_bench_ref.err; tail -c 800 gpurun_out/s14_bench_ref.json
timeout 200 python bench.py --workload poisson_mat_4096 --steps 50 --warmup 10 > gpurun_out/s14_bench_mat.json 2> gpurun_out/s14_bench_mat.err; tail -c 1500 gpurun_out/s14_bench_mat.json
timeout 600 bash profiles/collect.sh r01 > gpurun_out/s14_collect.log 2>&1; tail -5 gpurun_out/s14_collect.log
