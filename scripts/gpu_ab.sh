#!/usr/bin/env bash
# GPU visit: parity tests, smoke, then bench.py with the r02 launch-chain options toggled (A/B).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "${TMO:-900}" "$@" > "gpurun_out/$name.log" 2>&1; rc=$?; echo "exit $rc" >> "gpurun_out/$name.log"; tail -${TAILN:-6} "gpurun_out/$name.log"; return $rc; }
line() { python - "$1" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads([l for l in open(f'gpurun_out/{n}.log') if l.startswith('{')][-1])
    print(n, 'value', round(d['value'],1), 'ms/fwd', round(d['ms_per_unet_forward'],4), 'launches/fwd', d['launches']['per_unet_forward'], 'e2e', round(d['e2e']['value'],1))
except Exception as e:
    print(n, 'failed', e)
PY
}
if [ "${TESTS:-1}" = "1" ]; then
  TMO=1500 run pytest_gpu python -m pytest tests/ -x -q -m gpu -p no:cacheprovider
  run smoke python -c "import __graft_entry__ as g; g.smoke()"
fi
TAILN=1 run bench python bench.py --steps 2 --warmup 3; line bench
VARIANTS=${VARIANTS:-NS2VC_PIP=0 NS2VC_MERGE_FF=0}
for v in $VARIANTS; do
  n="bench_$(echo "$v" | tr '=,' '__')"
  env $(echo "$v" | tr ',' ' ') timeout 600 python bench.py --steps 2 --warmup 3 > "gpurun_out/$n.log" 2>&1; line "$n"
done
if [ "${SPAN:-0}" = "1" ]; then
  timeout 300 python scripts/span_trace.py > gpurun_out/span.log 2>&1; tail -8 gpurun_out/span.log
  for v in $VARIANTS; do
    n="span_$(echo "$v" | tr '=,' '__')"
    env $(echo "$v" | tr ',' ' ') timeout 300 python scripts/span_trace.py > "gpurun_out/$n.log" 2>&1; echo "--- $n"; tail -8 "gpurun_out/$n.log"
  done
fi
