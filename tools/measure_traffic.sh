#!/bin/bash
# One gpurun call: per-kernel DRAM bytes of 1 and 2 ADPM2 iterations at the bench workload -> gpurun_out/traffic_T{2,3}.csv
# (post-process here with tools/traffic_from_ncu.py).  Usage: tools/measure_traffic.sh [precision] [batch]
prec=${1:-tf32}; B=${2:-4096}
for T in 2 3; do
  MDT_GRAPH=0 timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    --csv --log-file gpurun_out/traffic_T$T.csv python tools/profile_step.py $prec $B $T > gpurun_out/traffic_T$T.log 2>&1
  tail -1 gpurun_out/traffic_T$T.log
done
