#!/bin/bash
# Round-2 evidence, per kernel: a handful of ncu metrics (duration, tensor-pipe activity, DRAM bytes / throughput,
# issue-slot use, registers) for EVERY launch of ~2 steps of the bench command, after a plain run of the same command
# that exited 0.  profiles/step_table.py then cuts out one whole step and writes the per-kernel-class table.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash profiles/collect_r2_step.sh'   (default workload: batch 48)
set -u
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eager-baseline"
timeout 300 $B > gpurun_out/r2s_plain.json 2> gpurun_out/r2s_plain.err || exit 1
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum
M=$M,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active
M=$M,launch__registers_per_thread,launch__grid_size,launch__block_size,l1tex__m_xbar2l1tex_read_bytes.sum
timeout 700 ncu --metrics $M --clock-control none --kernel-name-base demangled -s 1535 -c 450 --csv \
    --log-file gpurun_out/r2_ncu_step_metrics.csv $B > gpurun_out/r2s_ncu.log 2>&1
tail -2 gpurun_out/r2s_ncu.log
python profiles/step_table.py gpurun_out/r2_ncu_step_metrics.csv > gpurun_out/r2_kernel_table.md
head -40 gpurun_out/r2_kernel_table.md
