set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity_gaps.py -m gpu -x -q -k "policy or rollout" > gpurun_out/r2n_tests.txt 2>&1; tail -15 gpurun_out/r2n_tests.txt
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --e2e-steps 0 --rollout-steps 128 > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; tail -3 gpurun_out/r2n_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2n_bench.json')); print(d['value'], d['rollout']['value'], d['rollout']['frac_of_value'])"
