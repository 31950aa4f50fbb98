#!/bin/bash
# Reproduces the round-1 evidence under profiles/ (run on a B200 box through gpurun from the repo root):
#   /usr/local/graft/bin/gpurun --timeout 1800 -- 'bash profiles/collect.sh'
# Every ncu pass is preceded by the same command line run plainly (exit 0), as the profiling recipe requires.
set -u
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 900 python bench.py --steps 20 --warmup 3 --dump-ops gpurun_out/r1_ops.json > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r1_bench_reference.json 2>> gpurun_out/r1_bench.err
for b in 8 16 32 48; do
  timeout 300 python bench.py --steps 10 --warmup 3 --batch $b --no-cpu-baseline 2>/dev/null | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print(json.dumps({'batch_per_gpu': d['config']['batch_per_gpu'], 'value': d['value'], 'e2e': d['e2e']['value'], 'ms_per_step': d['ms_per_step'], 'clocks': d['clocks']}))"
done > gpurun_out/r1_batch_sweep.jsonl
timeout 300 $B > gpurun_out/plain1.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1535 -c 640 --csv \
    --log-file gpurun_out/r1_ncu_launches.csv $B > gpurun_out/ncu_launches.log 2>&1
# full captures: the heaviest single GEMM (g_a.2: 5x5 stride-2 conv + fused GDN = 2nd conv_gdn launch of a step) and the
# kernel with the largest share of the step (fused ResidualUnit tail = 4th conv_gdn launch)
timeout 300 $B > gpurun_out/plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_gdn_tc_kernel -s 1 -c 1 \
    -o gpurun_out/r1_prof_conv5x5_gdn $B > gpurun_out/ncu_top1.log 2>&1
timeout 300 $B > gpurun_out/plain3.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_gdn_tc_kernel -s 3 -c 1 \
    -o gpurun_out/r1_prof_ru_tail $B > gpurun_out/ncu_top2.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/r1_nvidia_smi.csv
tail -c 600 gpurun_out/r1_bench.json; echo; tail -2 gpurun_out/ncu_top1.log; tail -2 gpurun_out/ncu_top2.log
