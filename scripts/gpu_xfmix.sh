#!/usr/bin/env bash
# GPU visit: panel-mode channel limits (NS2VC_XF_MAXC1 / MAXC2 / MAXCP) - parity of each variant, then interleaved benches.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
line() { python - "$1" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads([l for l in open(f'gpurun_out/{n}.log') if l.startswith('{')][-1])
    print(n, 'value', round(d['value'],1), 'ms/fwd', round(d['ms_per_unet_forward'],4), 'launches/fwd', d['launches']['per_unet_forward'])
except Exception as e:
    print(n, 'failed', e)
PY
}
SEL="test_full_forward_matches_reference_fixture or test_full_forward_vs_oracle_shapes or test_tiny_forward_every_op"
i=0
for v in ${VARIANTS}; do
  i=$((i+1)); n="mix${i}_$(echo "$v" | tr '=,' '__')"
  if [ "${PARITY:-1}" = "1" ]; then
    env $(echo "$v" | tr ',' ' ') timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -p no:cacheprovider -k "$SEL" > "gpurun_out/par_$n.log" 2>&1; echo "parity $v: $(tail -1 gpurun_out/par_$n.log)"
  fi
done
i=0
for rep in 1 2; do
for v in ${VARIANTS}; do
  i=$((i+1)); n="mix${i}_$(echo "$v" | tr '=,' '__')"
  env $(echo "$v" | tr ',' ' ') timeout 600 python bench.py --steps 2 --warmup 3 > "gpurun_out/$n.log" 2>&1; line "$n"
done
done
