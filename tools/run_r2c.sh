# gpurun --timeout 1800 -- "bash tools/run_r2c.sh": parity, A/B of the kernel variants, one ncu --set full capture of a steady-state step
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_tests.txt 2>&1; tail -15 gpurun_out/r2c_tests.txt
AB_REPS=1 timeout 900 python tools/ab_libs.py > gpurun_out/r2c_ab.txt 2>&1; cat gpurun_out/r2c_ab.txt
timeout 300 python tools/ncu_step.py 200 4 > gpurun_out/r2c_plain.log 2>&1 && \
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name regex:^k_ --launch-skip 812 --launch-count 4 -f -o gpurun_out/r2c python tools/ncu_step.py 200 4 > gpurun_out/r2c_ncu.log 2>&1; tail -5 gpurun_out/r2c_ncu.log
