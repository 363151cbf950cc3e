# gpurun --gpus 2 --timeout 900 -- "bash tools/run_2gpu.sh": the driver's command at 2 GPUs (torchrun) and at 1 GPU on the same box
set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02b_bench_2gpu.json 2> gpurun_out/r02b_bench_2gpu.err; cut -c1-220 gpurun_out/r02b_bench_2gpu.json; tail -2 gpurun_out/r02b_bench_2gpu.err
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02b_bench_1gpu_same_box.json 2> gpurun_out/r02b_bench_1gpu_same_box.err; cut -c1-220 gpurun_out/r02b_bench_1gpu_same_box.json
python -c "
import json
for f in ('r02b_bench_2gpu', 'r02b_bench_1gpu_same_box'):
    d = json.load(open('gpurun_out/%s.json' % f)); print(f, d['value'], d['ms_per_step'], d['rollout']['value'], d['rollout']['frac_of_value'], d['roofline']['traffic'])
"
