#!/usr/bin/env bash
# ncu evidence for the condition encoders: launch list of ONE Pre_model.infer (cfg2 shape) and a full capture of its GEMMs
# (ENC instantiation: conv-FFN / masked residual epilogues; launches 21-23 = the first PhoneEncoder layer's FFN1, FFN2 and the
# next layer's out-projection at 8192 rows); then the whole-pipeline parity test.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "${TMO:-600}" "$@" > "gpurun_out/$name.log" 2>&1; rc=$?; echo "exit $rc" >> "gpurun_out/$name.log"; tail -${TAILN:-3} "gpurun_out/$name.log" | cut -c1-400; return $rc; }
TAILN=12 run pipeline python -m pytest tests/test_pipeline.py -q -m gpu -p no:cacheprovider
run ncu_pre_list ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/pre_launches.csv python scripts/ncu_pre_target.py
run ncu_pre_full ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:gemm_tc_kernel<\(int\)64, \(bool\)0, \(bool\)0, \(bool\)1>" -s 21 -c 3 -f -o gpurun_out/prof_pre python scripts/ncu_pre_target.py
ls -la gpurun_out/ | tail -8
