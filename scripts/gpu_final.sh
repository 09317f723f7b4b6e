#!/usr/bin/env bash
# Round-end GPU visit: the full -m gpu suite, smoke(), both bench arms, the in-step timelines, a memcheck pass and the ncu evidence
# (denoiser launch list, condition-encoder launch list + full capture); everything lands in gpurun_out/ (scripts/summarize_*.py
# turn it into profiles/).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; s=$(date +%s); timeout "${TMO:-900}" "$@" > "gpurun_out/$name.log" 2>&1; rc=$?; echo "exit $rc ($(( $(date +%s) - s )) s)" >> "gpurun_out/$name.log"; tail -${TAILN:-4} "gpurun_out/$name.log" | cut -c1-${CUT:-400}; return $rc; }
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
TMO=1800 run final_pytest_gpu python -m pytest tests/ -x -q -m gpu -p no:cacheprovider
run final_smoke python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
TAILN=2 CUT=200 run final_bench python bench.py
TAILN=2 CUT=200 run final_bench_ref python bench.py --impl reference --steps 1 --warmup 0
run final_span python scripts/span_trace.py
run final_trace python scripts/trace_gemm.py
TMO=900 run final_memcheck compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -p no:cacheprovider -k "test_full_forward_vs_oracle_shapes and 3-8-1"
TMO=900 run final_memcheck_pre compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_pre_model_gpu.py -x -q -m gpu -p no:cacheprovider -k "test_full_config_vs_oracle_shapes and 3-300-131"
if [ "${NCU:-1}" = "1" ]; then
  run ncu_list ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv python scripts/ncu_target.py 2
  run ncu_pre_list ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/pre_launches.csv python scripts/ncu_pre_target.py
  run ncu_pre_full ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:gemm_tc_kernel<\(int\)64, \(bool\)0, \(bool\)0, \(bool\)1>" -s 21 -c 3 -f -o gpurun_out/prof_pre python scripts/ncu_pre_target.py
fi
