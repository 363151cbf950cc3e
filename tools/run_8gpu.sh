# gpurun --gpus 8 --timeout 600 -- "bash tools/run_8gpu.sh": 8-GPU torchrun bench (weak scaling, 65 536 envs per GPU)
set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 200 --warmup 20 --e2e-steps 10 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; tail -c 1200 gpurun_out/bench_8gpu.json; tail -3 gpurun_out/bench_8gpu.err
