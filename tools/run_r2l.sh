# the other BASELINE configurations through bench.py on one GPU (smoke of --config before the multi-GPU runs)
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --config cfg4 --envs-per-gpu 131072 --steps 60 --warmup 10 --no-cpu-baseline --e2e-steps 10 > gpurun_out/r2l_cfg4.json 2> gpurun_out/r2l_cfg4.err; tail -3 gpurun_out/r2l_cfg4.err; cut -c1-400 gpurun_out/r2l_cfg4.json
timeout 900 python bench.py --config cfg5 --steps 40 --warmup 5 --no-cpu-baseline --e2e-steps 0 > gpurun_out/r2l_cfg5.json 2> gpurun_out/r2l_cfg5.err; tail -3 gpurun_out/r2l_cfg5.err; cut -c1-300 gpurun_out/r2l_cfg5.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2l_ref.json 2> gpurun_out/r2l_ref.err; cut -c1-300 gpurun_out/r2l_ref.json
