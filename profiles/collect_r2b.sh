#!/bin/bash
# ncu --set full (source-level) of single layers through tests/gpu_conv_bench.py: the ResidualUnit 1x1 head and the 1x1 tail form.
set -u
mkdir -p gpurun_out
cap() {  # name, case, kernel regex
  REPS=2 timeout 200 python tests/gpu_conv_bench.py $2 > gpurun_out/r2b_plain_$1.log 2>&1 || return
  REPS=2 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$3" -s 2 -c 1 \
      -o gpurun_out/r2b_prof_$1 python tests/gpu_conv_bench.py $2 > gpurun_out/r2b_ncu_$1.log 2>&1
  ncu -i gpurun_out/r2b_prof_$1.ncu-rep --page raw --csv > gpurun_out/r2b_prof_$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/r2b_prof_$1.ncu-rep --page source --csv > gpurun_out/r2b_prof_$1_source.csv 2>/dev/null
  rm -f gpurun_out/r2b_prof_$1.ncu-rep
  tail -1 gpurun_out/r2b_plain_$1.log
}
cap ruhead 1x1_ru_192_96 conv_tc_kernel
cap qkv 1x1_qkv_192_576 conv_tc_kernel
cap resgelu 1x1_rutail_96_192_resgelu conv_tc_kernel
