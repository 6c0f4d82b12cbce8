#!/bin/bash
# A/B of two builds of the library on ONE box (box-to-box variance is ~2 %): tools/ab_libs.sh <variant.so> [rounds]
# alternates the variant with the current lib/liberfnet_b200.so and prints training / inference throughput per run.
set -e
cd "$(dirname "$0")/.."
V=$1; R=${2:-2}
cp erfnet_pytorch_b200/lib/liberfnet_b200.so /tmp/cur.so
for i in $(seq $R); do
  for name in cur variant; do
    if [ $name = cur ]; then cp /tmp/cur.so erfnet_pytorch_b200/lib/liberfnet_b200.so; else cp "$V" erfnet_pytorch_b200/lib/liberfnet_b200.so; fi
    ERF_BENCH_EXTRA=0 timeout 300 python bench.py 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$name', round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), round(d['fwd']['images_per_s'],1))"
  done
done
cp /tmp/cur.so erfnet_pytorch_b200/lib/liberfnet_b200.so
