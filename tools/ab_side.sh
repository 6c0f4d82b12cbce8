#!/bin/bash
# same-box sweep of the side-stream weight gradients with SM caps: tools/ab_side.sh "main,side" ...
run() { env "$@" ERF_BENCH_EXTRA=0 timeout 300 python bench.py --steps 20 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$*', round(d['value'],1), round(d['ms_per_step'],4))"; }
run ERF_WGRAD_SIDE=0
for cap in "$@"; do run ERF_WGRAD_SIDE=1 ERF_SIDE_CAP=$cap; done
run ERF_WGRAD_SIDE=0
