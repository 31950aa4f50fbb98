#!/bin/bash
# Final evidence of round 2 (one B200, default workload = batch 48 x 512 x 768): `ncu --set full` of the dominant kernel
# (-> profiles/r2_roofline_traffic.json, which bench.py reports as `roofline.traffic`), the bench line with per-launch op
# dump and clocks sampled during the run, per-launch ncu metrics of one step and the per-kernel / per-layer tables.
# Every ncu pass follows a plain run of the same command that exited 0.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash profiles/collect_r2_final.sh'
set -u
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eager-baseline"
timeout 300 $B > gpurun_out/r2_plain.json 2> gpurun_out/r2_plain.err || exit 1
KEY=$(python -c "
import json
d=json.loads(open('gpurun_out/r2_plain.json').read().strip().splitlines()[-1])['roofline']
print(d['kernel'] + ' ' + d['layer'])")
echo "dominant kernel: $KEY"
if [ "${SKIP_FULL:-0}" != "1" ]; then  # (SKIP_FULL=1: keep the committed capture of the unchanged dominant kernel)
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:ru_pair_tc_kernel" -s 2 -c 1 \
    -o gpurun_out/r2_prof_ru_pair $B > gpurun_out/r2_ncu_ru_pair.log 2>&1
ncu -i gpurun_out/r2_prof_ru_pair.ncu-rep --page raw --csv > gpurun_out/r2_prof_ru_pair_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_prof_ru_pair.ncu-rep --page source --csv > gpurun_out/r2_prof_ru_pair_source.csv 2>/dev/null
rm -f gpurun_out/r2_prof_ru_pair.ncu-rep
python profiles/make_traffic_json.py gpurun_out/r2_prof_ru_pair_raw.csv "$KEY" profiles/r2_roofline_traffic.json > /dev/null
cp profiles/r2_roofline_traffic.json gpurun_out/r2_roofline_traffic.json
fi
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap \
    --format=csv -lms 200 > gpurun_out/r2_final_clocks.csv &
SMI=$!
timeout 600 python bench.py --steps 20 --warmup 5 --dump-ops gpurun_out/r2_ops_final.json > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err || { kill $SMI; exit 1; }
kill $SMI
bash profiles/collect_r2_step.sh > gpurun_out/r2_step.log 2>&1
(cd profiles && python layer_table.py ../gpurun_out/r2_ncu_step_metrics.csv ../gpurun_out/r2_ops_final.json > ../gpurun_out/r2_layer_table.md)
head -14 gpurun_out/r2_kernel_table.md
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_final.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_compact']['value'], d['launches_per_step'], d['clocks'], d['roofline']['frac'], d['roofline']['whole_step']['frac'], d['roofline']['traffic'], d['gpu_eager_baseline']['variants'], d['cpu_baseline']['value'], d['config']['batch_per_gpu'])"
