cd /root/repo
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; s=$(date +%s); timeout "${TMO:-900}" "$@" > "gpurun_out/$name.log" 2>&1; rc=$?; echo "exit $rc ($(( $(date +%s) - s )) s)" >> "gpurun_out/$name.log"; tail -${TAILN:-3} "gpurun_out/$name.log" | cut -c1-500; return $rc; }
TAILN=15 TMO=1500 run pytest_gpu python -m pytest tests/ -q -m gpu -p no:cacheprovider --durations=8
run ncu_pre_full ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:gemm_tc_kernel<\(int\)64, \(bool\)0, \(bool\)0, \(bool\)1>" -s 12 -c 5 -f -o gpurun_out/prof_pre python scripts/ncu_pre_target.py
ls -la gpurun_out | tail -5
