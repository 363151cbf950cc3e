# new policy kernel (v3), breakdown, bench with rollout, fresh ncu capture of one step of the current build
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity_gaps.py -m gpu -x -q -k "policy or rollout" > gpurun_out/r2r_policy_tests.txt 2>&1; tail -5 gpurun_out/r2r_policy_tests.txt
timeout 300 python tools/rollout_breakdown.py > gpurun_out/r2r_breakdown.txt 2>&1; tail -3 gpurun_out/r2r_breakdown.txt
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --e2e-steps 0 --rollout-steps 128 > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; tail -3 gpurun_out/r2r_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2r_bench.json')); print(d['value'], d['ms_per_step'], d['rollout']['value'], d['rollout']['frac_of_value'])"
timeout 300 python tools/ncu_step.py 200 4 > gpurun_out/r2r_plain2.log 2>&1 && \
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name regex:^k_ --launch-skip 609 --launch-count 3 -f -o gpurun_out/r2r_step python tools/ncu_step.py 200 4 > gpurun_out/r2r_ncu2.log 2>&1; tail -3 gpurun_out/r2r_ncu2.log
