set -x
mkdir -p gpurun_out
AB_REPS=2 timeout 900 python tools/ab_libs.py > gpurun_out/r2t_ab.txt 2>&1; cat gpurun_out/r2t_ab.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_gaps.py -m gpu -x -q > gpurun_out/r2t_tests.txt 2>&1; tail -3 gpurun_out/r2t_tests.txt
