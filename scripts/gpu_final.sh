#!/usr/bin/env bash
# Round-end GPU visit: the full -m gpu suite, smoke(), both bench arms, the in-step timelines and a memcheck pass; everything
# lands in gpurun_out/ (copy what is to be judged into profiles/).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "${TMO:-900}" "$@" > "gpurun_out/$name.log" 2>&1; rc=$?; echo "exit $rc" >> "gpurun_out/$name.log"; tail -${TAILN:-4} "gpurun_out/$name.log" | cut -c1-400; return $rc; }
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
TMO=1800 run final_pytest_gpu python -m pytest tests/ -x -q -m gpu -p no:cacheprovider
run final_smoke python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
TAILN=1 run final_bench python bench.py
TAILN=1 run final_bench_ref python bench.py --impl reference --steps 1 --warmup 0
run final_span python scripts/span_trace.py
run final_trace python scripts/trace_gemm.py
TMO=900 run final_memcheck compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -p no:cacheprovider -k "test_full_forward_vs_oracle_shapes and 3-8-1"
