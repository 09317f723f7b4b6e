#!/usr/bin/env bash
# One GPU-box visit: staged parity tests (separate processes so a trapped kernel cannot poison the
# rest), smoke, short bench.  Logs land in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout "${TMO:-600}" "$@" > "gpurun_out/$name.log" 2>&1; echo "exit $?" >> "gpurun_out/$name.log"; tail -${TAILN:-15} "gpurun_out/$name.log"; }
run t1_simt python -m pytest tests/test_gpu_parity.py -q -k "simt or in_tree" -p no:cacheprovider
run t2_steps python -m pytest tests/test_gpu_parity.py -q -k "kernel_bit_exact" -p no:cacheprovider
run t3_tc python -m pytest tests/test_gpu_parity.py -q -k "tiny_forward_every_op and tc" -p no:cacheprovider
TMO=1200 run t4_rest python -m pytest tests/test_gpu_parity.py -q -k "not tiny_forward_every_op and not kernel_bit_exact and not in_tree" -p no:cacheprovider
run smoke python -c "import __graft_entry__ as g; g.smoke()"
TMO=900 run bench python bench.py --steps ${BENCH_STEPS:-1} --warmup 3
