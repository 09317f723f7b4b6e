#!/usr/bin/env bash
# GPU visit for the condition encoders: their parity tests, smoke(), a subset of the denoiser tests (the GEMM kernel gained an
# instantiation), one bench line (carries the `pre_model` key).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "${TMO:-600}" "$@" > "gpurun_out/$name.log" 2>&1; rc=$?; echo "exit $rc" >> "gpurun_out/$name.log"; tail -${TAILN:-15} "gpurun_out/$name.log" | cut -c1-${CUT:-600}; return $rc; }
TAILN=40 run pre_tests python -m pytest tests/test_pre_model_gpu.py -q -p no:cacheprovider
TAILN=4 run smoke python -c "import __graft_entry__ as g; g.smoke()"
if [ "${SUBSET:-1}" = "1" ]; then TAILN=6 run unet_subset python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -x -k "tiny_forward_every_op or full_forward_matches_reference or kernel_bit_exact"; fi
if [ "${BENCH:-1}" = "1" ]; then TAILN=2 CUT=6000 run bench python bench.py --steps 2 --warmup 3; fi
