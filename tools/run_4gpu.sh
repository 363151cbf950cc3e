set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02b_bench_4gpu.json 2> gpurun_out/r02b_bench_4gpu.err; tail -1 gpurun_out/r02b_bench_4gpu.err
python -c "
import json
d = json.load(open('gpurun_out/r02b_bench_4gpu.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['rollout']['value'], d['rollout']['frac_of_value'], d['e2e']['value'], d['e2e'].get('frac_of_d2h_ceiling'))
"
