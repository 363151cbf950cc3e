# ncu evidence of the final build: launch list of the driver's bench command, then one --set full capture of a steady-state step
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 0 --rollout-steps 0 > gpurun_out/r02b_plain.json 2> gpurun_out/r02b_plain.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:^k_ --launch-skip 510 --launch-count 60 --csv --log-file gpurun_out/r02b_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 0 --rollout-steps 0 > gpurun_out/r02b_ncu1.log 2>&1
timeout 300 python tools/ncu_step.py 200 4 > gpurun_out/r02b_plain2.log 2>&1 && \
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name regex:^k_ --launch-skip 609 --launch-count 3 -f -o gpurun_out/r02b_final python tools/ncu_step.py 200 4 > gpurun_out/r02b_ncu2.log 2>&1; tail -2 gpurun_out/r02b_ncu2.log
