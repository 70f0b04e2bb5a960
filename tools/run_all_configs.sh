#!/bin/bash
# One gpurun call: bench.py for every BASELINE.json config (1 GPU, default precision) -> gpurun_out/r02_bench_<cfg>.json
mkdir -p gpurun_out
timeout 900 python bench.py --config cfg2 > gpurun_out/r02_bench_cfg2.json 2> gpurun_out/r02_bench_cfg2.err; cut -c1-300 gpurun_out/r02_bench_cfg2.json
timeout 300 python bench.py --config cfg1 --steps 10 --warmup 5 --also "" > gpurun_out/r02_bench_cfg1.json 2> gpurun_out/r02_bench_cfg1.err; cut -c1-300 gpurun_out/r02_bench_cfg1.json
timeout 900 python bench.py --config cfg3 --e2e-steps 2 --also "" > gpurun_out/r02_bench_cfg3.json 2> gpurun_out/r02_bench_cfg3.err; cut -c1-300 gpurun_out/r02_bench_cfg3.json
timeout 1500 python bench.py --config cfg4 --e2e-steps 2 --also "" > gpurun_out/r02_bench_cfg4.json 2> gpurun_out/r02_bench_cfg4.err; cut -c1-300 gpurun_out/r02_bench_cfg4.json
timeout 1500 python bench.py --config cfg5 --e2e-steps 2 --also "" > gpurun_out/r02_bench_cfg5.json 2> gpurun_out/r02_bench_cfg5.err; cut -c1-300 gpurun_out/r02_bench_cfg5.json
