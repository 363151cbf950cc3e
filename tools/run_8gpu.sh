# gpurun --gpus 8 --timeout 1500 -- "bash tools/run_8gpu.sh": every BASELINE configuration on the 8 GPUs of one box + the box's D2H ceiling
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29501 bench.py --gpus 8 --steps 200 --warmup 20 > gpurun_out/r02b_bench_8gpu.json 2> gpurun_out/r02b_bench_8gpu.err; tail -2 gpurun_out/r02b_bench_8gpu.err; cut -c1-250 gpurun_out/r02b_bench_8gpu.json
timeout 400 $TR --master-port 29502 bench.py --gpus 8 --config cfg4 --steps 100 --warmup 10 --e2e-steps 10 > gpurun_out/r02b_bench_8gpu_cfg4.json 2> gpurun_out/r02b_bench_8gpu_cfg4.err; tail -2 gpurun_out/r02b_bench_8gpu_cfg4.err; cut -c1-250 gpurun_out/r02b_bench_8gpu_cfg4.json
timeout 400 $TR --master-port 29503 bench.py --gpus 8 --config cfg5 --steps 50 --warmup 5 --e2e-steps 0 > gpurun_out/r02b_bench_8gpu_cfg5.json 2> gpurun_out/r02b_bench_8gpu_cfg5.err; tail -2 gpurun_out/r02b_bench_8gpu_cfg5.err; cut -c1-250 gpurun_out/r02b_bench_8gpu_cfg5.json
timeout 200 $TR --master-port 29504 tools/d2h_ceiling.py --gpus 8 > gpurun_out/r02b_d2h_ceiling_8gpu.json 2> gpurun_out/r02b_d2h_ceiling_8gpu.err; cat gpurun_out/r02b_d2h_ceiling_8gpu.json
nvidia-smi topo -m > gpurun_out/r02b_topo_8gpu.txt 2>&1; head -14 gpurun_out/r02b_topo_8gpu.txt
