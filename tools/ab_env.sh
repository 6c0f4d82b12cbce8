set -x
timeout 300 python -m pytest tests/test_gpu_conv1d_tc.py -x -q 2>&1 | tail -5
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
ERF_WGRAD_BULK=0 python tools/bench_wgrad.py > gpurun_out/wgrad_bulk0.txt 2>&1; cat gpurun_out/wgrad_bulk0.txt
ERF_WGRAD_BULK=1 python tools/bench_wgrad.py > gpurun_out/wgrad_bulk1.txt 2>&1; cat gpurun_out/wgrad_bulk1.txt
for i in 1 2; do for b in 0 1; do
ERF_WGRAD_BULK=$b ERF_BENCH_EXTRA=0 timeout 300 python bench.py --steps 20 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('bulk=$b', round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), round(d['fwd']['images_per_s'],1))"
done; done
