# gpurun --timeout 1800 -- "bash tools/run_r2d.sh": parity, bench, pipelined-handles diagnostic
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_tests.txt 2>&1; tail -15 gpurun_out/r2d_tests.txt
timeout 600 python tools/pipeline_ab.py > gpurun_out/r2d_pipe.txt 2>&1; cat gpurun_out/r2d_pipe.txt
timeout 300 python tools/ab_libs.py 'tools/none*.so' > gpurun_out/r2d_ab.txt 2>&1; cat gpurun_out/r2d_ab.txt
timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; tail -3 gpurun_out/r2d_bench.err; cut -c1-300 gpurun_out/r2d_bench.json
