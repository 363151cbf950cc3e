# refresh of the 1-GPU bench lines with the final build (tests first)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_tests.txt 2>&1; tail -3 gpurun_out/r02b_tests.txt
timeout 900 python bench.py --steps 200 --warmup 20 > gpurun_out/r02b_bench_1gpu.json 2> gpurun_out/r02b_bench_1gpu.err; tail -2 gpurun_out/r02b_bench_1gpu.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02b_bench_driver_cmd.json 2> gpurun_out/r02b_bench_driver_cmd.err
timeout 300 python tools/rollout_breakdown.py > gpurun_out/r02b_rollout_breakdown.txt 2>&1; tail -1 gpurun_out/r02b_rollout_breakdown.txt
python -c "
import json
for f in ('r02b_bench_1gpu', 'r02b_bench_driver_cmd'):
    d = json.load(open('gpurun_out/%s.json' % f)); print(f, d['value'], d['ms_per_step'], d['rollout']['value'], d['rollout']['frac_of_value'], d['roofline']['traffic'], d['roofline']['frac'], d['e2e']['value'])
"
