# end-of-session check on the GPU box: GPU tests, smoke, both bench arms, the mat workload
set -x
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py > gpurun_out/final_bench_wave.json 2> gpurun_out/final_bench_wave.err; tail -c 700 gpurun_out/final_bench_wave.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; tail -c 300 gpurun_out/final_bench_ref.json
timeout 200 python bench.py --workload poisson_mat_4096 --steps 50 --warmup 10 > gpurun_out/final_bench_mat.json 2> gpurun_out/final_bench_mat.err; tail -c 900 gpurun_out/final_bench_mat.json
