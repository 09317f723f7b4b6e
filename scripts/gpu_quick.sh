#!/usr/bin/env bash
# Quick GPU visit: the driver's own pytest command, smoke, one bench line, in-situ span timeline.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "${TMO:-900}" "$@" > "gpurun_out/$name.log" 2>&1; rc=$?; echo "exit $rc" >> "gpurun_out/$name.log"; tail -${TAILN:-6} "gpurun_out/$name.log"; return $rc; }
TMO=1500 run pytest_gpu python -m pytest tests/ -x -q -m gpu -p no:cacheprovider
run smoke python -c "import __graft_entry__ as g; g.smoke()"
TAILN=1 run bench python bench.py --steps ${BENCH_STEPS:-2} --warmup 3
TAILN=${SPAN_TAIL:-8} run span python scripts/span_trace.py
