#!/bin/bash
# Throughput versus internal chunk size (MDT_MAX_BATCH) and external batch.
for mb in 1024 2048; do
  MDT_MAX_BATCH=$mb timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --also "" --e2e-steps 1 2>/dev/null > /tmp/o.json
  python -c "import json; d=json.load(open('/tmp/o.json')); print('max_batch', $mb, 'batch 4096:', round(d['value'],1))"
done
MDT_MAX_BATCH=8192 timeout 300 python bench.py --batch 8192 --steps 2 --warmup 3 --no-cpu --also "" --e2e-steps 1 2>/dev/null > /tmp/o.json
python -c "import json; d=json.load(open('/tmp/o.json')); print('max_batch 8192 batch 8192:', round(d['value'],1))"
