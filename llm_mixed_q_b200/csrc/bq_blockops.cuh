// bq_blockops.cuh — register-level block quantisation shared by the fused kernels (attention softmax / epilogue,
// quantising GEMM epilogues, LayerNorm+quantize).  Element arithmetic comes from bq_numerics.cuh (bit-identical to the
// reference's torch emulation); this header only adds the "16 values of one block live in one thread" plumbing.
#pragma once
#include <cuda_bf16.h>

#include "bq_numerics.cuh"

namespace bq {

__device__ __forceinline__ float u2f(uint32_t u) { return __uint_as_float(u); }
__device__ __forceinline__ uint32_t f2u(float f) { return __float_as_uint(f); }
// upper halves of two fp32 -> one bf16x2 word (exact for values with <= 8 significant bits)
__device__ __forceinline__ uint32_t pack_bf16_trunc(float lo, float hi) { return __byte_perm(f2u(lo), f2u(hi), 0x7632); }
__device__ __forceinline__ uint32_t pack_bf16_rn(float lo, float hi) {
  __nv_bfloat162 t2 = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t2);
}

// cold path: one element through the literal (reference-order) arithmetic; kept out of line, scalars only
template <int KIND>
__device__ __noinline__ float quant_literal_1(float x, uint32_t mbits, FmtParams p) {
  const BlockState st = block_state<KIND>(__uint_as_float(mbits), p);
  return quant_elem<KIND>(x, st, p);
}

// one element given the block's max |x| bits (already substituted: never 0)
template <int KIND>
__device__ __forceinline__ float quant_with_max(float x, uint32_t mbits, const FmtParams& p) {
  const FastState fs = fast_state<KIND>(mbits, p);
  return fs.ok ? quant_elem_fast<KIND>(x, fs, p) : quant_literal_1<KIND>(x, mbits, p);
}

// Quantise 16 consecutive SIGNED values held by one thread (one reference block) in place.
template <int KIND>
__device__ __forceinline__ void quantize_signed16(float (&v)[16], const FmtParams& p) {
  uint32_t m = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) m = max(m, f2u(v[i]) & 0x7fffffffu);
  if (m == 0) {                       // all-zero block: every element passes through; -0.0 -> +0.0 like the reference's blend
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = 0.f;
    return;
  }
  const FastState fs = fast_state<KIND>(m, p);
  if (fs.ok) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = quant_elem_fast<KIND>(v[i], fs, p);
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = quant_literal_1<KIND>(v[i], m, p);
  }
}

// runtime-kind wrapper for the two block formats whose all-zero blocks need no tensor-global information
__device__ __forceinline__ void quantize_signed16_rt(float (&v)[16], const FmtParams& p) {
  if (p.kind == kBlockFP) quantize_signed16<kBlockFP>(v, p);
  else quantize_signed16<kBlockMinifloat>(v, p);
}

}  // namespace bq
