#!/bin/bash
# One gpurun call (round 2): ncu launch list of one ADPM2 iteration, DRAM traffic of 1 and 2 iterations, ncu --set full of the
# dominant kernels.  Everything lands in gpurun_out/ (summaries are copied to profiles/ by hand after reading them here).
mkdir -p gpurun_out
MDT_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r02_launches_tf32.csv python tools/profile_step.py tf32 4096 2 > gpurun_out/ncu_l.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches_tf32.csv > gpurun_out/r02_launches_tf32.summary.txt 2>&1; head -14 gpurun_out/r02_launches_tf32.summary.txt
tools/measure_traffic.sh tf32 4096
cap() {  # name regex skip count
  MDT_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c $4 -o gpurun_out/r02_prof_$1 -f python tools/profile_step.py tf32 4096 2 > gpurun_out/ncu_$1.log 2>&1
  ls -la gpurun_out/r02_prof_$1.ncu-rep 2>&1 | cut -c30-
}
cap attn_frag_self 'attn_frag_kernel<\(int\)1, \(int\)0>' 4 3
cap attn_frag_cross 'attn_frag_kernel<\(int\)1, \(int\)16>' 2 2
cap attn_frag_cross_l2 'attn_frag_kernel<\(int\)1, \(int\)4>' 2 2
cap ff_chain 'ff_chain_kernel' 2 2
cap gemm_tma 'gemm_tma_kernel' 20 12
cap resnet_small 'resnet_small_kernel' 0 2
cap step_update 'step_update_kernel' 0 2
