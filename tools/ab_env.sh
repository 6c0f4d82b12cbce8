#!/bin/bash
# A/B of an environment switch on ONE box: tools/ab_env.sh VAR [rounds] -- alternates VAR=0 / VAR=1 and prints training / e2e / inference throughput per run
V=$1; R=${2:-2}
for i in $(seq $R); do for b in 0 1; do
env $V=$b ERF_BENCH_EXTRA=0 timeout 300 python bench.py --steps 20 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$V=$b', round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), round(d['fwd']['images_per_s'],1))"
done; done
