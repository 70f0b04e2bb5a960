#!/bin/bash
# One gpurun call: full GPU test suite, default bench (+ bf16), reference arm, ncu launch list, ncu --set full of the packed cross-attention launches.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; cat gpurun_out/bench_default.json
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference.json 2>/dev/null; cat gpurun_out/bench_reference.json
MDT_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_tf32.csv python tools/profile_step.py tf32 4096 2 > gpurun_out/ncu1.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_tf32.csv 2>/dev/null | head -12
MDT_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:gemm_attn_kernel<\(int\)1, \(int\)3>' -s 4 -c 3 -o gpurun_out/prof_gemm_attn_packed_cross -f python tools/profile_step.py tf32 4096 2 > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out/prof_gemm_attn_packed_cross.ncu-rep; tail -2 gpurun_out/ncu2.log
