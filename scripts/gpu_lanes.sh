#!/usr/bin/env bash
# Re-measure the multi-stream batch lanes on the current build (graph replay + persistent GEMM + attention v2).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt
for L in 1 2 4 8; do
  NS2VC_LANES=$L timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_lanes$L.log 2>&1
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_lanes$L.log').read().strip().splitlines()[-1])
    print('lanes', $L, 'value', round(d['value'],1), 'ms/fwd', round(d['ms_per_unet_forward'],4), 'e2e', round(d['e2e']['value'],1))
except Exception as e:
    print('lanes', $L, 'failed', e)
PY
done
