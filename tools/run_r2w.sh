set -x
mkdir -p gpurun_out
AB_REPS=2 timeout 900 python tools/ab_libs.py > gpurun_out/r2w_ab.txt 2>&1; cat gpurun_out/r2w_ab.txt
