#!/bin/bash
# A/B of an environment switch on the bench workload:  tools/ab.sh VAR [A B]  (runs VAR=A and VAR=B twice, default 0 and 1, default precision, no CPU baseline)
v=$1
a=${2:-0}; b=${3:-1}
for x in $a $b $a $b; do
  env $v=$x timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 0 --also '' 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    l = l.strip()
    if l.startswith('{'):
        d = json.loads(l); print('$v=$x', d['value'], d['ms_per_step'])"
done
