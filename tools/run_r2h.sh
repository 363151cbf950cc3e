set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_tests.txt 2>&1; tail -8 gpurun_out/r2h_tests.txt
timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; tail -3 gpurun_out/r2h_bench.err; cut -c1-300 gpurun_out/r2h_bench.json
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h_smoke.txt 2>&1; tail -2 gpurun_out/r2h_smoke.txt
