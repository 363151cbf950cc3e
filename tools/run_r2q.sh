# A/B: compact out-of-line tracker sums (code size 309 -> 164 KB); ncu --set full of the new policy kernel
set -x
mkdir -p gpurun_out
AB_REPS=2 timeout 900 python tools/ab_libs.py > gpurun_out/r2q_ab.txt 2>&1; cat gpurun_out/r2q_ab.txt
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name regex:k_policy --launch-skip 300 --launch-count 2 -f -o gpurun_out/r2q_policy python tools/rollout_breakdown.py > gpurun_out/r2q_ncu.log 2>&1; tail -2 gpurun_out/r2q_ncu.log
