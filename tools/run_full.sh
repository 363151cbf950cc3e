# gpurun --timeout 1800 -- "bash tools/run_full.sh": tests, smoke, bench (both arms), ncu launch list + full capture, sweeps -> gpurun_out/
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/t10.txt 2>&1; tail -3 gpurun_out/t10.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke10.txt 2>&1; tail -1 gpurun_out/smoke10.txt
python bench.py --steps 300 --warmup 30 > gpurun_out/bench10.json 2> gpurun_out/bench10.err; tail -c 300 gpurun_out/bench10.json; tail -2 gpurun_out/bench10.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench10_ref.json 2>> gpurun_out/bench10.err; cut -c1-200 gpurun_out/bench10_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 300 --launch-count 150 --csv --log-file gpurun_out/r01e_launches.csv python bench.py --steps 150 --warmup 20 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu10a.log 2>&1
ncu --set full --clock-control none --import-source on --launch-skip 300 --launch-count 3 -f -o gpurun_out/r01e python bench.py --steps 150 --warmup 20 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu10b.log 2>&1
python tools/sweep_f.py > gpurun_out/sweeps10.txt 2>&1; tail -6 gpurun_out/sweeps10.txt
