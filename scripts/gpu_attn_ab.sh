#!/usr/bin/env bash
# A/B of the attention variants on ONE box (cross-box noise is +-3-4 %): parity first, then bench lines.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "${TMO:-600}" "$@" > "gpurun_out/$name.log" 2>&1; rc=$?; echo "exit $rc" >> "gpurun_out/$name.log"; tail -${TAILN:-8} "gpurun_out/$name.log"; return $rc; }
K="full_forward or deterministic"
ok_nat=0; ok_wide=0
TMO=900 run p_v2 python -m pytest tests/test_gpu_parity.py -q -x -k "$K" -p no:cacheprovider && ok_nat=1
NS2VC_ATTN_PB=128 TMO=900 run p_v2_pb128 python -m pytest tests/test_gpu_parity.py -q -x -k "$K" -p no:cacheprovider && ok_wide=1
echo "parity: natural=$ok_nat pb128=$ok_wide"
[ $ok_nat = 1 ] && TAILN=1 TMO=900 run b_v2 python bench.py --steps 1 --warmup 3
[ $ok_wide = 1 ] && NS2VC_ATTN_PB=128 TAILN=1 TMO=900 run b_v2_pb128 python bench.py --steps 1 --warmup 3
NS2VC_ATTN=v1 TAILN=1 TMO=900 run b_v1 python bench.py --steps 1 --warmup 3
if [ $ok_nat = 1 ]; then TAILN=60 run span python scripts/span_trace.py; elif [ $ok_wide = 1 ]; then NS2VC_ATTN_PB=128 TAILN=60 run span python scripts/span_trace.py; fi
