set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,sm__inst_executed_pipe_tensor.sum --clock-control none --kernel-name regex:k_policy --launch-skip 300 --launch-count 4 --csv --log-file gpurun_out/r2o_policy_ncu.csv python tools/rollout_breakdown.py > gpurun_out/r2o_ncu.log 2>&1; tail -2 gpurun_out/r2o_ncu.log; tail -30 gpurun_out/r2o_policy_ncu.csv | cut -c1-60,200-400
