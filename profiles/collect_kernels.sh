#!/bin/bash
# Per-kernel `ncu --set full` captures for the kernel classes named in BASELINE.json's north_star (run after
# profiles/collect.sh; the plain command has exited 0 there):
#   /usr/local/graft/bin/gpurun --timeout 1800 -- 'bash profiles/collect_kernels.sh'
set -u
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $B > gpurun_out/plain_k.log 2>&1 || exit 1
cap() {  # name, kernel regex, skip, count
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c $4 \
      -o gpurun_out/r1_prof_$1 $B > gpurun_out/ncu_$1.log 2>&1
  tail -1 gpurun_out/ncu_$1.log
}
cap conv_gdn1 'conv_gdn_tc_kernel<\(int\)1>' 0 3        # conv + GDN (first conv, 5x5 s2 192->192, ...)
cap conv_tc_none 'conv_tc_kernel<\(int\)0, \(bool\)1>' 0 6   # EPI_NONE launches: h_a tail, slice-loop partial-sum convs (N = 224, M2)
cap attn 'win_attn_tc_kernel<\(int\)8' 0 1
# (gc_forward_kernel / eb_forward_kernel: cap gc 'gc_forward_kernel' 2 1; cap eb 'eb_forward_kernel' 0 1)
