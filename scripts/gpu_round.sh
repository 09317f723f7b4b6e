#!/usr/bin/env bash
# One GPU-box visit: staged parity tests (separate processes so a trapped kernel cannot poison the
# rest), smoke, short bench, ncu launch list + one full capture.  Logs land in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "${TMO:-600}" "$@" > "gpurun_out/$name.log" 2>&1; echo "exit $?" >> "gpurun_out/$name.log"; tail -${TAILN:-12} "gpurun_out/$name.log"; }
run t1_tiny python -m pytest tests/test_gpu_parity.py -q -k "tiny_forward_every_op or in_tree or kernel_bit_exact" -p no:cacheprovider
TMO=1200 run t2_rest python -m pytest tests/test_gpu_parity.py -q -k "not tiny_forward_every_op and not kernel_bit_exact and not in_tree" -p no:cacheprovider
run smoke python -c "import __graft_entry__ as g; g.smoke()"
TMO=900 TAILN=3 run bench python bench.py --steps ${BENCH_STEPS:-1} --warmup 3
if [ "${AB:-0}" = "1" ]; then for L in 1 4 8; do NS2VC_LANES=$L TMO=900 TAILN=1 run bench_lanes$L python bench.py --steps 1 --warmup 3; done; fi
if [ "${NCU:-1}" = "1" ]; then
  TMO=600 TAILN=3 run ncu_list ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv python scripts/ncu_target.py 2
  TMO=900 TAILN=3 run ncu_full ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 100 -c 4 -f -o gpurun_out/prof_gemm python scripts/ncu_target.py 1
fi
