#!/usr/bin/env bash
# Builds libbq_b200.so (sm_100a only) in-tree: llm_mixed_q_b200/libbq_b200.so
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${HERE}/../libbq_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -Xcompiler -fvisibility=hidden)
mkdir -p "${HERE}/build"
pids=()
for f in common quantize gemm_sm100 gemm_xform_sm100 ops attention_sm100 loss peer softmax_quant; do
  "${NVCC}" "${FLAGS[@]}" ${BQ_EXTRA_FLAGS:-} ${BQ_PTXAS_V:+-Xptxas -v} -c "${HERE}/${f}.cu" -o "${HERE}/build/${f}.o" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
"${NVCC}" -shared -o "${OUT}" "${HERE}"/build/{common,quantize,gemm_sm100,gemm_xform_sm100,ops,attention_sm100,loss,peer,softmax_quant}.o -lcudart_static -ldl -lrt -lpthread
echo "built ${OUT}"
