set -x
mkdir -p gpurun_out
timeout 300 python tools/debug_trace.py gazebo_list_regimes_random_frames_seed3 > gpurun_out/r2g_debug.txt 2>&1; tail -12 gpurun_out/r2g_debug.txt
AB_REPS=2 timeout 900 python tools/ab_libs.py > gpurun_out/r2g_ab.txt 2>&1; cat gpurun_out/r2g_ab.txt
