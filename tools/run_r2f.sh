set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_tests.txt 2>&1; tail -15 gpurun_out/r2f_tests.txt
AB_REPS=2 timeout 900 python tools/ab_libs.py > gpurun_out/r2f_ab.txt 2>&1; cat gpurun_out/r2f_ab.txt
