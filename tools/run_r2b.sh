# gpurun --timeout 1500 -- "bash tools/run_r2b.sh": GPU parity suites + bench of the current build
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_tests.txt 2>&1; tail -15 gpurun_out/r2b_tests.txt
timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -3 gpurun_out/r2b_bench.err; cut -c1-300 gpurun_out/r2b_bench.json
