#!/bin/bash
# Round-2 evidence: launch list of the bench command and `ncu --set full` captures (source-level) of the fused
# ResidualUnit kernel.  Every ncu pass follows a plain run of the same command that exited 0.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash profiles/collect_r2.sh'
set -u
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eager-baseline"
timeout 300 $B > gpurun_out/r2_plain.json 2> gpurun_out/r2_plain.err || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -s 1535 -c 640 --csv \
    --log-file gpurun_out/r2_ncu_launches.csv $B > gpurun_out/r2_ncu_launches.log 2>&1
cap() {  # name, kernel regex, skip, count
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c $4 \
      -o gpurun_out/r2_prof_$1 $B > gpurun_out/r2_ncu_$1.log 2>&1
  tail -1 gpurun_out/r2_ncu_$1.log
  ncu -i gpurun_out/r2_prof_$1.ncu-rep --page raw --csv > gpurun_out/r2_prof_$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/r2_prof_$1.ncu-rep --page source --csv > gpurun_out/r2_prof_$1_source.csv 2>/dev/null
}
cap ru_pair 'ru_pair_tc_kernel' 2 1
python profiles/make_traffic_json.py gpurun_out/r2_prof_ru_pair_raw.csv "ru_pair_tc_kernel M=589824 N=96 K=864 3x3s1 tailN=192" gpurun_out/r2_roofline_traffic.json
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_throttle_reasons.active --format=csv > gpurun_out/r2_nvidia_smi.csv
