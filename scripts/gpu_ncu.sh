#!/usr/bin/env bash
# ncu evidence for profiles/: launch list with DRAM bytes, one full capture of the GEMM and of the attention kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "${TMO:-900}" "$@" > "gpurun_out/$name.log" 2>&1; rc=$?; echo "exit $rc" >> "gpurun_out/$name.log"; tail -${TAILN:-3} "gpurun_out/$name.log"; return $rc; }
run ncu_list ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv python scripts/ncu_target.py 2
run ncu_gemm ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 60 -c 6 -f -o gpurun_out/prof_gemm python scripts/ncu_target.py 1
run ncu_xf ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:gemm_tc_kernel<\(int\)64, \(bool\)0, \(bool\)1>" -s 4 -c 4 -f -o gpurun_out/prof_xf python scripts/ncu_target.py 1
run ncu_attn ncu --set full --clock-control none --import-source on -k regex:attn_v2 -s 0 -c 6 -f -o gpurun_out/prof_attn python scripts/ncu_target.py 1
