#!/bin/bash
# Final evidence of round 2 (one B200): bench line with per-launch op dump, per-launch ncu metrics of one step,
# `ncu --set full` of the attention core through tests/gpu_attn_bench.py.  Every ncu pass follows a plain run of the
# same command that exited 0.
#   /usr/local/graft/bin/gpurun --timeout 1200 -- 'bash profiles/collect_r2_final.sh'
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap \
    --format=csv -lms 200 > gpurun_out/r2_final_clocks.csv &
SMI=$!
timeout 400 python bench.py --steps 20 --warmup 5 --dump-ops gpurun_out/r2_ops_final.json > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err || { kill $SMI; exit 1; }
kill $SMI
bash profiles/collect_r2_step.sh > gpurun_out/r2_step.log 2>&1
(cd profiles && python layer_table.py ../gpurun_out/r2_ncu_step_metrics.csv ../gpurun_out/r2_ops_final.json > ../gpurun_out/r2_layer_table.md)
timeout 120 python tests/gpu_attn_bench.py > gpurun_out/r2_attn_bench.log 2>&1 || exit 1
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:win_attn_tc_kernel<\(int\)8" -s 3 -c 1 \
    -o gpurun_out/r2_prof_attn python tests/gpu_attn_bench.py > gpurun_out/r2_ncu_attn.log 2>&1
ncu -i gpurun_out/r2_prof_attn.ncu-rep --page raw --csv > gpurun_out/r2_prof_attn_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_prof_attn.ncu-rep --page source --csv > gpurun_out/r2_prof_attn_source.csv 2>/dev/null
rm -f gpurun_out/r2_prof_attn.ncu-rep
tail -3 gpurun_out/r2_attn_bench.log
head -12 gpurun_out/r2_kernel_table.md
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_final.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_compact']['value'], d['launches_per_step'], d['clocks'], d['roofline']['frac'], d['roofline']['whole_step']['frac'], d['gpu_eager_baseline']['best_images_per_s'], d['cpu_baseline']['value'])"
