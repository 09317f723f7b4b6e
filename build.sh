#!/usr/bin/env bash
# Build the C-ABI shared library for sm_100a (in-tree, travels to the GPU box with gpurun).
set -euo pipefail
cd "$(dirname "$0")"
OUT=ns2vc_b200/_C
mkdir -p "$OUT"
SRC="ns2vc_b200/csrc/kernels_misc.cu ns2vc_b200/csrc/gemm_simt.cu ns2vc_b200/csrc/gemm_tc.cu ns2vc_b200/csrc/attention.cu ns2vc_b200/csrc/attention_v2.cu ns2vc_b200/csrc/engine.cu ns2vc_b200/csrc/pre_kernels.cu ns2vc_b200/csrc/pre_engine.cu"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared \
     ${NVCC_EXTRA:-} -o "$OUT/libns2vc_b200.so" $SRC
echo "built $OUT/libns2vc_b200.so"
