# last check of the in-tree build: GPU tests, smoke, the driver's bench command and the reference arm
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/final3_tests.txt 2>&1; tail -3 gpurun_out/final3_tests.txt
timeout 120 python -c "
import __graft_entry__ as g; g.smoke()" > gpurun_out/final3_smoke.txt 2>&1; tail -1 gpurun_out/final3_smoke.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/final3_bench.json 2> gpurun_out/final3_bench.err; cut -c1-260 gpurun_out/final3_bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final3_ref.json 2>/dev/null; cut -c1-200 gpurun_out/final3_ref.json
