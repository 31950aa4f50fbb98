#!/bin/bash
# Reproduces the round-1 evidence under profiles/ (run on a B200 box through gpurun from the repo root):
#   /usr/local/graft/bin/gpurun --timeout 1800 -- 'bash profiles/collect.sh'
# Every ncu pass is preceded by the same command line run plainly (exit 0), as the profiling recipe requires.
set -u
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 900 python bench.py --steps 20 --warmup 3 --dump-ops gpurun_out/r1_ops.json > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r1_bench_reference.json 2>> gpurun_out/r1_bench.err
timeout 300 $B > gpurun_out/plain1.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1535 -c 640 --csv \
    --log-file gpurun_out/r1_ncu_launches.csv $B > gpurun_out/ncu_launches.log 2>&1
timeout 300 $B > gpurun_out/plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 788 -c 1 \
    -o gpurun_out/r1_prof_top $B > gpurun_out/ncu_top.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/r1_nvidia_smi.csv
tail -c 600 gpurun_out/r1_bench.json; echo; tail -2 gpurun_out/ncu_top.log
