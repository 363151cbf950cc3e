# final 1-GPU evidence of the last build: tests, smoke, bench lines (2 000 settle steps), reference arm, launch list, ncu --set full of one step
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02f_tests.txt 2>&1; tail -3 gpurun_out/r02f_tests.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02f_smoke.txt 2>&1; tail -1 gpurun_out/r02f_smoke.txt
timeout 900 python bench.py --steps 200 --warmup 20 > gpurun_out/r02f_bench_1gpu.json 2> gpurun_out/r02f_bench_1gpu.err; tail -2 gpurun_out/r02f_bench_1gpu.err; cut -c1-300 gpurun_out/r02f_bench_1gpu.json
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02f_bench_driver_cmd.json 2> gpurun_out/r02f_bench_driver_cmd.err; cut -c1-200 gpurun_out/r02f_bench_driver_cmd.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02f_bench_reference_arm.json 2> gpurun_out/r02f_ref.err; cut -c1-300 gpurun_out/r02f_bench_reference_arm.json
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 0 --rollout-steps 0 --settle 150 > gpurun_out/r02f_plain.json 2> gpurun_out/r02f_plain.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:^k_ --launch-skip 510 --launch-count 60 --csv --log-file gpurun_out/r02f_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 0 --rollout-steps 0 --settle 150 > gpurun_out/r02f_ncu1.log 2>&1
timeout 300 python tools/ncu_step.py 200 4 > gpurun_out/r02f_plain2.log 2>&1 && \
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name regex:^k_ --launch-skip 609 --launch-count 3 -f -o gpurun_out/r02f_final python tools/ncu_step.py 200 4 > gpurun_out/r02f_ncu2.log 2>&1; tail -3 gpurun_out/r02f_ncu2.log
timeout 600 python bench.py --config cfg2 --steps 100 --warmup 10 --no-cpu-baseline --e2e-steps 0 --rollout-steps 0 --settle 150 > gpurun_out/r02f_bench_1gpu_cfg2.json 2> gpurun_out/r02f_cfg2.err; cut -c1-200 gpurun_out/r02f_bench_1gpu_cfg2.json
timeout 300 python tools/rollout_breakdown.py > gpurun_out/r02f_rollout_breakdown.txt 2>&1; tail -1 gpurun_out/r02f_rollout_breakdown.txt
