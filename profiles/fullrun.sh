# end-of-session check on the GPU box: GPU tests, smoke, both bench arms
set -x
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 400 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 400 gpurun_out/final_bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; tail -c 600 gpurun_out/final_bench_ref.json
