# gpurun --gpus 8 --timeout 900 -- "bash tools/run_8gpu_cfg3.sh": cfg3 on the 8 GPUs of one box with the final build (200-step line and the driver's command)
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29501 bench.py --gpus 8 --steps 200 --warmup 20 > gpurun_out/r02b_bench_8gpu.json 2> gpurun_out/r02b_bench_8gpu.err; tail -2 gpurun_out/r02b_bench_8gpu.err
timeout 400 $TR --master-port 29502 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02b_bench_8gpu_driver_cmd.json 2> gpurun_out/r02b_bench_8gpu_driver_cmd.err
python -c "
import json
for f in ('r02b_bench_8gpu', 'r02b_bench_8gpu_driver_cmd'):
    d = json.load(open('gpurun_out/%s.json' % f)); print(f, d['value'], d['ms_per_step'], d['rollout']['value'], d['rollout']['frac_of_value'], d['e2e']['value'], d['e2e'].get('frac_of_d2h_ceiling'), d['config']['stats_allreduces_in_timed_region'])
"
