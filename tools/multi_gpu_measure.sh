#!/bin/bash
# Multi-GPU evidence under `gpurun --gpus N`:  tools/multi_gpu_measure.sh "<sweep gpu counts>" "<bench gpu counts>" [test]
#   (a) optional: the 2-rank token-equality test, (b) the cfg-5 sweep checksum at the given GPU counts (same rows, same seed: identical
#   tokens expected at every count), (c) bench.py weak scaling at the given counts.  Appends to gpurun_out/r02_multigpu.log.
mkdir -p gpurun_out
log=gpurun_out/r02_multigpu_$(nvidia-smi -L | wc -l)gpu.log; : > $log
echo "visible GPUs: $(nvidia-smi -L | wc -l)" | tee -a $log
if [ -n "$3" ]; then timeout 900 python -m pytest tests/test_gpu_parity.py -q -k two_rank 2>&1 | tail -2 | tee -a $log; fi
for n in $1; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n)) \
    tools/sweep.py --rows 32768 --chunk 4096 --cond-scale 5 --timesteps 64 --seed 4 2>/dev/null | grep rows= | tee -a $log
done
for n in $2; do
  if [ $n = 1 ]; then cmd="python bench.py"; else cmd="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29800 + n)) bench.py"; fi
  timeout 900 $cmd --gpus $n --steps 3 --warmup 3 --e2e-steps 2 --also "" --no-cpu > gpurun_out/r02_bench_cfg2_${n}gpu.json 2> gpurun_out/r02_bench_cfg2_${n}gpu.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/r02_bench_cfg2_${n}gpu.json').read().strip().splitlines()[-1]); print('bench gpus', d['n_gpus'], 'samples/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],1))" | tee -a $log
done
