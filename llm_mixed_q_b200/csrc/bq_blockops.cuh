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

// block_log (block_log.py:23-69, log.py:22-56) for the 16-BIT CARRIERS of the fused kernels.  Two things the reference does cannot be
// reproduced block-locally / in a bf16 carrier, and both only concern outputs below 2^-126 ("block_log carrier rule"):
//   * an ALL-ZERO block takes the tensor-global minimum g of the non-zero block maxima (block_log.py:50-53); its elements come out as
//     +2^(ceil(log2 g) - 127) < 2^-126 * g.  Here such a block stays 0.
//   * elements far below their block maximum clamp to 2^emin with emin = ceil(log2 blockmax) - 127, which is below the smallest
//     normal number whenever blockmax <= 1 (and below bf16's smallest denormal 2^-133 for blockmax <= 2^-7).  Here every output that
//     the reference puts below 2^-126 is 0 or 2^-126.
// Every output >= 2^-126 is the reference's, bit for bit (same delta = 0.1 * 2^emin added before log2, same rint(log2f) through the
// exponent shortcut with the libdevice fallback next to the sqrt(2) cliff).  |deviation| <= 2^-126 per element, stated and tested
// (tests/test_gpu_block_log_fused.py); a GEMM output moves by at most K * 2^-126 * max|other operand|.
static __device__ __noinline__ void quantize_blocklog16_cold(float* v, uint32_t mbits, FmtParams p) {
  const BlockState st = block_state<kBlockLog>(__uint_as_float(mbits), p);
  for (int i = 0; i < 16; ++i) v[i] = quant_elem<kBlockLog>(v[i], st, p);
}
__device__ __forceinline__ void quantize_signed16_blocklog(float (&v)[16], const FmtParams& p) {
  uint32_t m = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) m = max(m, f2u(v[i]) & 0x7fffffffu);
  if (m == 0) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = 0.f;
    return;
  }
  bool fast = p.fast_fmt && m < 0x7f800000u && m >= 0x00800000u;           // finite, normal block maximum
  int i0 = 0, i1 = 0;
  if (fast) {
    int b = p.eb_top_i - ceil_log2_i(__uint_as_float(m));
    b = min(max(b, 0), p.bias_hi_i);
    i0 = -b;
    i1 = p.eb_top_i - b;
    fast = i1 <= 127 && i1 >= i0 && i1 >= -125;
  }
  if (!fast) {                                                             // inf / NaN / denormal maxima, exotic ranges: literal arithmetic
    float t[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) t[i] = v[i];
    quantize_blocklog16_cold(t, m, p);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = t[i];
    return;
  }
  const float delta = __fmul_rn(i0 >= -126 ? pow2_i(i0) : pow2_t((float)i0), 0.1f);       // log.py:49-52 (denormal or 0 below 2^-126)
  const int lo = i0 + 127, hi = i1 + 127;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float w = __fadd_rn(v[i], delta);
    const float a = __fadd_rn(fabsf(v[i]), delta);
    uint32_t unused = 0xffffffffu;
    const int eb = min(max(rint_log2_biased_f<false>(a, unused), lo), hi);
    v[i] = (w == 0.f || eb <= 0) ? 0.f : copysignf(__int_as_float(eb << 23), w);
  }
}

template <>
__device__ __forceinline__ void quantize_signed16<kBlockLog>(float (&v)[16], const FmtParams& p) { quantize_signed16_blocklog(v, p); }

// silu(g) * u (Llama MLP: act_fn(gate_proj(h)) * up_proj(h), reference modeling_llama.py:246) in torch-CUDA's op order:
// silu(g) = g / (1 + expf(-g)) (ActivationSiluKernel.cu), then one multiply.  Shared by the streaming silu*mul quantizer
// (quantize.cu) and the gated GEMM epilogue (gemm_sm100.cu) so that both give the same bits.
__device__ __forceinline__ float silu_mul1(float g, float u) {
  return __fmul_rn(__fdiv_rn(g, __fadd_rn(1.0f, expf(-g))), u);
}

// runtime-kind wrapper for the block formats of the fused kernels (block_log: carrier rule above)
__device__ __forceinline__ void quantize_signed16_rt(float (&v)[16], const FmtParams& p) {
  if (p.kind == kBlockFP) quantize_signed16<kBlockFP>(v, p);
  else if (p.kind == kBlockLog) quantize_signed16_blocklog(v, p);
  else quantize_signed16<kBlockMinifloat>(v, p);
}

}  // namespace bq
