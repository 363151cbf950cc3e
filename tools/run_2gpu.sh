# gpurun --gpus 2 --timeout 900 -- "bash tools/run_2gpu.sh": 2-GPU torchrun bench and the 1-GPU bench on the same box
set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 300 --warmup 30 > gpurun_out/bench7_2gpu.json 2> gpurun_out/bench7_2gpu.err; tail -c 1500 gpurun_out/bench7_2gpu.json; tail -3 gpurun_out/bench7_2gpu.err
python bench.py --gpus 1 --steps 300 --warmup 30 --no-cpu-baseline > gpurun_out/bench7_1gpu.json 2> gpurun_out/bench7_1gpu.err; python tools/pick_bench.py < gpurun_out/bench7_1gpu.json
