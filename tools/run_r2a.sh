# gpurun --timeout 1500 -- "bash tools/run_r2a.sh": GPU parity suites, then the bench line of the unchanged round-1 kernels at steady state
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.txt 2>&1; tail -6 gpurun_out/r2a_tests.txt
timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -3 gpurun_out/r2a_bench.err; cut -c1-600 gpurun_out/r2a_bench.json
timeout 300 python bench.py --config cfg2 --steps 200 --no-cpu-baseline --rollout-steps 0 > gpurun_out/r2a_bench_cfg2.json 2> gpurun_out/r2a_bench_cfg2.err; tail -3 gpurun_out/r2a_bench_cfg2.err; cut -c1-300 gpurun_out/r2a_bench_cfg2.json
