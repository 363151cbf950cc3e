# A/B: wide predicated scans / warp-cooperative window walk; then the GPU parity tests on the wide8coop build; new policy kernel
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity_gaps.py -m gpu -x -q -k "policy or rollout" > gpurun_out/r2p_policy_tests.txt 2>&1; tail -15 gpurun_out/r2p_policy_tests.txt
timeout 300 python tools/rollout_breakdown.py > gpurun_out/r2p_breakdown.txt 2>&1; tail -3 gpurun_out/r2p_breakdown.txt
timeout 900 python tools/ab_libs.py > gpurun_out/r2p_ab.txt 2>&1; cat gpurun_out/r2p_ab.txt
cp tools/libftl_wide8coop.so continiousenvironment_follower_leader_b200/csrc/libftl.so
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_gaps.py -m gpu -x -q > gpurun_out/r2p_tests.txt 2>&1; tail -5 gpurun_out/r2p_tests.txt
