# gpurun --timeout 1500 -- "bash tools/run_ab.sh": GPU parity tests, then same-box A/B timing of csrc/libftl.so against every tools/libftl_*.so (tools/build_variant.py)
set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu.txt 2>&1; tail -4 gpurun_out/tests_gpu.txt
AB_REPS=2 timeout 600 python tools/ab_libs.py > gpurun_out/ab.txt 2>&1; cat gpurun_out/ab.txt
