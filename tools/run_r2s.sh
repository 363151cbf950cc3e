# A/B: records in shared memory, walk prefetch, window prefetch; then the GPU parity tests on the default build
set -x
mkdir -p gpurun_out
AB_REPS=2 timeout 900 python tools/ab_libs.py > gpurun_out/r2s_ab.txt 2>&1; cat gpurun_out/r2s_ab.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_gaps.py -m gpu -x -q > gpurun_out/r2s_tests.txt 2>&1; tail -5 gpurun_out/r2s_tests.txt
