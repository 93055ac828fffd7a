// softmax_quant.cu — the attention glue of the formats the one-kernel attention (attention_sm100.cu) does not serve: block_log, whose
// matmuls leave the y operand UNQUANTISED in fp32 (reference quantized_functions/matmul.py:286-297), and any head_dim / mask the
// fused kernel rejects.  Three kernels replace the reference's op-by-op passes over the S x S scores:
//
//   softmax_quant_kernel       scores (fp32) -> [/ sqrt(d)] -> + causal / key-padding mask -> max(finfo.min) -> softmax -> x-quantizer of
//                              matmul_1 / bmm_1 -> bf16 P   (models/llama_quantized/modeling_llama.py:309-337,
//                              models/opt_quantized/modeling_opt.py:246-312, bert_quantized/modeling_bert.py:366-435).  The reference
//                              streams the B*h*S*S tensor ~10 times in fp32 (mask add, max, softmax, the quantizer's ~45 passes); here it
//                              is read once (only the causally visible part) and the quantised probabilities are written once in bf16.
//   rope_split_kernel          Llama rotary embedding of q followed by matmul_0's x-quantizer (head-major bf16) and of k followed by the
//                              error-free split of the UNQUANTISED fp32 result into three bf16 planes (k = k0 + k1 + k2 exactly: the
//                              operand format of the fp32-equivalent tensor-core GEMM, bq_bmm_split_tn).
//   split3_transposed_kernel   v (fp32, token-major) -> three bf16 planes of v^T per head ([d][S], keys contiguous): the K-major B
//                              operand of P @ V.
//
// Products of a power-of-two (block_log) x operand with the planes of y are exact, so QK^T and PV differ from the reference's fp32
// matmuls by accumulation order only.
#include <cuda_bf16.h>
#include <math.h>

#include "bq_blockops.cuh"
#include "bq_internal.h"

namespace bq {
namespace {

__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream2(void* p, uint32_t a, uint32_t b) {
  asm volatile("st.global.cs.v2.b32 [%0], {%1,%2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ float warp_max(float v) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_add(float v) {
  for (int o = 16; o > 0; o >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// exp(s - m): libdevice expf like torch (precise mode), or one ex2.approx of the fp32 product (s - m) * log2(e) — relative error
// <= ~(2 + 1.44 |s - m|) ulp, which moves a probability only when it sits that close to a rounding boundary of its format; the same
// default and the same switch (bq_set_attention_precise_exp) as the one-kernel attention (stated deviation 3)
template <bool FAST>
__device__ __forceinline__ float exp_sm(float s, float m) {
  const float t = __fsub_rn(s, m);
  if (!FAST) return expf(t);
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(__fmul_rn(t, 1.4426950408889634f)));
  return r;
}

struct SmqArgs {
  const float* s;             // [batch][Sq][lds]
  __nv_bfloat16* p;           // [batch][Sq][ldp]
  int64_t lds, ss, ldp, sp;   // row / batch strides (elements)
  int batch, Sq, Sk, heads;
  float mul;                  // scores * mul (1 / sqrt(d)) when scale != 0 — torch-CUDA evaluates tensor / python_float that way
  int scale;
  int causal;                 // key j takes part in query row i iff j <= i (Sq == Sk)
  const uint32_t* kmask;      // [batch / heads][kwords] key-validity bits, or nullptr
  int kwords;
  FmtParams f;
};

constexpr float kNegMax = -3.4028234663852886e38f;      // torch.finfo(float32).min: what the reference's masks hold

// quantise 4 consecutive probabilities of a block whose maximum (over its 16 elements = 4 lanes) has bits `m`
template <int KIND>
__device__ __forceinline__ float4 quant4_block(float4 v, uint32_t m, const FmtParams& f) {
  if (m == 0) return make_float4(0.f, 0.f, 0.f, 0.f);
  if (KIND == kBlockLog) {
    // carrier rule of bq_blockops.cuh (quantize_signed16_blocklog): block-local, outputs below 2^-126 are 0 or 2^-126
    float t[4] = {v.x, v.y, v.z, v.w};
    bool fast = f.fast_fmt && m < 0x7f800000u && m >= 0x00800000u;
    int i0 = 0, i1 = 0;
    if (fast) {
      int b = f.eb_top_i - ceil_log2_i(__uint_as_float(m));
      b = min(max(b, 0), f.bias_hi_i);
      i0 = -b;
      i1 = f.eb_top_i - b;
      fast = i1 <= 127 && i1 >= i0 && i1 >= -125;
    }
    if (!fast) {
      const BlockState st = block_state<kBlockLog>(__uint_as_float(m), f);
#pragma unroll 1
      for (int i = 0; i < 4; ++i) t[i] = quant_elem<kBlockLog>(t[i], st, f);
    } else {
      const float delta = __fmul_rn(i0 >= -126 ? pow2_i(i0) : pow2_t((float)i0), 0.1f);
      const int lo = i0 + 127, hi = i1 + 127;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float w = __fadd_rn(t[i], delta), a = __fadd_rn(fabsf(t[i]), delta);
        uint32_t unused = 0xffffffffu;
        const int eb = min(max(rint_log2_biased_f<false>(a, unused), lo), hi);
        t[i] = (w == 0.f || eb <= 0) ? 0.f : copysignf(__int_as_float(eb << 23), w);
      }
    }
    return make_float4(t[0], t[1], t[2], t[3]);
  } else {
    const FastState fs = fast_state<KIND>(m, f);
    if (fs.ok)
      return make_float4(quant_elem_fast<KIND>(v.x, fs, f), quant_elem_fast<KIND>(v.y, fs, f), quant_elem_fast<KIND>(v.z, fs, f),
                         quant_elem_fast<KIND>(v.w, fs, f));
    return make_float4(quant_literal_1<KIND>(v.x, m, f), quant_literal_1<KIND>(v.y, m, f), quant_literal_1<KIND>(v.z, m, f),
                       quant_literal_1<KIND>(v.w, m, f));
  }
}

// One query row per warp; lane l owns the float4 chunks c = j * 32 + l (keys 4c .. 4c+3), so every load / store instruction of the
// warp covers 512 / 256 contiguous bytes and a block of 16 keys is held by 4 neighbouring lanes (two shuffles for its maximum).
// NV = chunks per lane held in registers (Sk <= 128 * NV); NV == 0: any Sk, three passes over the row (the 2nd and 3rd from L1 / L2).
template <int KIND, int NV, bool FAST>
__global__ void __launch_bounds__(256) softmax_quant_kernel(SmqArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t rows = (int64_t)a.batch * a.Sq;
  const int nchunk = a.Sk >> 2;
  for (int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < rows; row += warps) {
    const int bi = (int)(row / a.Sq), qi = (int)(row - (int64_t)bi * a.Sq);
    const float* sr = a.s + (int64_t)bi * a.ss + (int64_t)qi * a.lds;
    __nv_bfloat16* pr = a.p + (int64_t)bi * a.sp + (int64_t)qi * a.ldp;
    const uint32_t* km = a.kmask ? a.kmask + (int64_t)(bi / a.heads) * a.kwords : nullptr;
    const int kvis = a.causal ? qi + 1 : a.Sk;                 // keys [0, kvis) pass the causal mask
    // causal == 2: the consumer (bq_bmm_split_tn, causal 2) reads keys up to the end of the 256-row block of its tile at most
    const int zchunk = a.causal == 2 ? min(nchunk, ((qi >> 8) + 1) << 6) : nchunk;
    // masked scores are finfo.min like the reference's additive mask + clamp leave them (a fully masked row is then uniform)
    auto load = [&](int c) {
      float4 v = make_float4(kNegMax, kNegMax, kNegMax, kNegMax);
      const int k0 = c << 2;
      if (c < nchunk && k0 < kvis) {
        v = ld_stream4(sr + k0);
        if (a.scale) { v.x = __fmul_rn(v.x, a.mul); v.y = __fmul_rn(v.y, a.mul); v.z = __fmul_rn(v.z, a.mul); v.w = __fmul_rn(v.w, a.mul); }
        uint32_t bits = 0xfu;
        if (km) bits = (km[k0 >> 5] >> (k0 & 31)) & 0xfu;
        if (k0 + 3 >= kvis) bits &= (1u << (kvis - k0)) - 1u;
        v.x = (bits & 1u) ? fmaxf(v.x, kNegMax) : kNegMax;
        v.y = (bits & 2u) ? fmaxf(v.y, kNegMax) : kNegMax;
        v.z = (bits & 4u) ? fmaxf(v.z, kNegMax) : kNegMax;
        v.w = (bits & 8u) ? fmaxf(v.w, kNegMax) : kNegMax;
      }
      return v;
    };
    // p = e * (1 / l): one correctly rounded reciprocal per row instead of an IEEE division per element (torch divides; the
    // product can differ from the quotient in the last bit, which moves a probability only when it sits on a rounding boundary —
    // same statement as for the row sum's order)
    auto finish = [&](float4 e, float inv_l, int c) {      // normalise, quantise the chunk inside its block of 16, store
      float4 p = make_float4(__fmul_rn(e.x, inv_l), __fmul_rn(e.y, inv_l), __fmul_rn(e.z, inv_l), __fmul_rn(e.w, inv_l));
      uint32_t m = max(max(f2u(p.x), f2u(p.y)), max(f2u(p.z), f2u(p.w)));        // probabilities are >= 0
      m = max(m, __shfl_xor_sync(0xffffffffu, m, 1));
      m = max(m, __shfl_xor_sync(0xffffffffu, m, 2));
      if (c < nchunk) {
        const float4 q = quant4_block<KIND>(p, m, a.f);
        st_stream2(pr + (c << 2), pack_bf16_rn(q.x, q.y), pack_bf16_rn(q.z, q.w));
      }
    };
    if (NV > 0) {
      float4 v[NV > 0 ? NV : 1];
      float m = kNegMax;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        v[j] = load(j * 32 + lane);
        m = fmaxf(m, fmaxf(fmaxf(v[j].x, v[j].y), fmaxf(v[j].z, v[j].w)));
      }
      m = warp_max(m);
      // A row with at least one visible key has m > finfo.min and every masked score contributes expf(finfo.min - m) == 0 exactly:
      // chunks that lie wholly behind the causal diagonal need no arithmetic (half of all chunks).  A fully masked row (m ==
      // finfo.min: every score equal) is uniform over ALL keys, as in the reference — then nothing is skipped.
      const int jvis = (m == kNegMax) ? NV : min(NV, (kvis + 127) >> 7);        // warp-uniform
      float l = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        if (j < jvis) {
          const bool in = (j * 32 + lane) < nchunk;
          v[j].x = in ? exp_sm<FAST>(v[j].x, m) : 0.f; v[j].y = in ? exp_sm<FAST>(v[j].y, m) : 0.f;
          v[j].z = in ? exp_sm<FAST>(v[j].z, m) : 0.f; v[j].w = in ? exp_sm<FAST>(v[j].w, m) : 0.f;
          l = __fadd_rn(l, __fadd_rn(__fadd_rn(v[j].x, v[j].y), __fadd_rn(v[j].z, v[j].w)));
        }
      }
      l = __frcp_rn(warp_add(l));
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        if (j < jvis) {
          finish(v[j], l, j * 32 + lane);
        } else if (j * 32 + lane < zchunk) {
          st_stream2(pr + ((j * 32 + lane) << 2), 0u, 0u);              // probabilities behind the diagonal: exact zeros
        }
      }
    } else {
      const int nj = (nchunk + 31) >> 5;
      float m = kNegMax;
      for (int j = 0; j < nj; ++j) {
        const float4 v = load(j * 32 + lane);
        m = fmaxf(m, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
      }
      m = warp_max(m);
      float l = 0.f;
      for (int j = 0; j < nj; ++j) {
        if (j * 32 + lane < nchunk) {
          const float4 v = load(j * 32 + lane);
          l = __fadd_rn(l, __fadd_rn(__fadd_rn(exp_sm<FAST>(v.x, m), exp_sm<FAST>(v.y, m)),
                                     __fadd_rn(exp_sm<FAST>(v.z, m), exp_sm<FAST>(v.w, m))));
        }
      }
      l = __frcp_rn(warp_add(l));
      for (int j = 0; j < nj; ++j) {
        const int c = j * 32 + lane;
        float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < nchunk) {
          const float4 v = load(c);
          e = make_float4(exp_sm<FAST>(v.x, m), exp_sm<FAST>(v.y, m), exp_sm<FAST>(v.z, m), exp_sm<FAST>(v.w, m));
        }
        finish(e, l, c);
      }
    }
  }
}

// ---- v2 (rows of up to 8192 keys): the row lives in SHARED memory, loops are rolled.
// v1 above keeps a 2048-key row in 64 registers per lane and unrolls everything: 118 registers (16 warps per SM), 6700 static
// instructions (107 KB of code: 2 of every 10 issue slots lost to instruction fetch) and ~11 divergence-capable branches per chunk from
// the inlined checked quantizer — 60 executed instructions per visible score, 0.69 ms per Llama-7B layer (profiles/r02_ncu_softmax_quant_v1.json).
// Here: the visible part of the row arrives by cp.async (no registers), three short rolled passes (max / exp + sum / normalise +
// quantise) read and write it in place, the block_log quantiser of a PROBABILITY (>= 0: no sign handling) is branch-free — rint(log2)
// from the exponent field with the distance to the sqrt(2) cliff min-accumulated and tested once per chunk.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ float4 lds4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts4(uint32_t a, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// 4 probabilities of one block (maximum bits m, shared by the block's 4 lanes) -> quantised values
template <int KIND>
__device__ __forceinline__ float4 quant4_probs(float4 p, uint32_t m, const FmtParams& f) {
  if (KIND != kBlockLog) return quant4_block<KIND>(p, m, f);
  if (m == 0) return make_float4(0.f, 0.f, 0.f, 0.f);
  if (!(f.fast_fmt && m < 0x7f800000u && m >= 0x00800000u)) return quant4_block<KIND>(p, m, f);
  int b = f.eb_top_i - ceil_log2_i(__uint_as_float(m));
  b = min(max(b, 0), f.bias_hi_i);
  const int i0 = -b, i1 = f.eb_top_i - b;
  if (i1 > 127 || i1 < -125) return quant4_block<KIND>(p, m, f);
  // delta = 0.1 * 2^emin (log.py:49-52): 2^emin is a normal number, a denormal (emin >= -149) or 0
  const uint32_t p2 = i0 >= -126 ? (uint32_t)(i0 + 127) << 23 : (i0 >= -149 ? 1u << (i0 + 149) : 0u);
  const float delta = __fmul_rn(__uint_as_float(p2), 0.1f);
  const int lo = max(i0 + 127, 0), hi = i1 + 127;          // biased exponents; 0 = "below 2^-126": flushed (carrier rule)
  uint32_t zacc = 0xffffffffu;
  const int e0 = min(max(rint_log2_biased_f<true>(__fadd_rn(p.x, delta), zacc), lo), hi);
  const int e1 = min(max(rint_log2_biased_f<true>(__fadd_rn(p.y, delta), zacc), lo), hi);
  const int e2 = min(max(rint_log2_biased_f<true>(__fadd_rn(p.z, delta), zacc), lo), hi);
  const int e3 = min(max(rint_log2_biased_f<true>(__fadd_rn(p.w, delta), zacc), lo), hi);
  if (zone_hit<kBlockLog>(zacc)) return quant4_block<KIND>(p, m, f);          // an element within 2^-13 of the sqrt(2) cliff: checked path
  return make_float4(__int_as_float(e0 << 23), __int_as_float(e1 << 23), __int_as_float(e2 << 23), __int_as_float(e3 << 23));
}

template <int KIND, bool FAST>
__global__ void __launch_bounds__(256) softmax_quant_smem_kernel(SmqArgs a, int warps_per_cta) {
  extern __shared__ __align__(16) uint8_t smq_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t rowb = (uint32_t)__cvta_generic_to_shared(smq_smem) + (uint32_t)warp * (uint32_t)a.Sk * 4u;
  const int64_t nw = (int64_t)gridDim.x * warps_per_cta;
  const int64_t rows = (int64_t)a.batch * a.Sq;
  const int nchunk = a.Sk >> 2;
  for (int64_t row = (int64_t)blockIdx.x * warps_per_cta + warp; row < rows; row += nw) {
    const int bi = (int)(row / a.Sq), qi = (int)(row - (int64_t)bi * a.Sq);
    const float* sr = a.s + (int64_t)bi * a.ss + (int64_t)qi * a.lds;
    __nv_bfloat16* pr = a.p + (int64_t)bi * a.sp + (int64_t)qi * a.ldp;
    const uint32_t* km = a.kmask ? a.kmask + (int64_t)(bi / a.heads) * a.kwords : nullptr;
    const int kvis = a.causal ? qi + 1 : a.Sk;
    const int zchunk = a.causal == 2 ? min(nchunk, ((qi >> 8) + 1) << 6) : nchunk;
    const int vchunk = min(nchunk, (kvis + 3) >> 2);              // chunks that hold a visible key
    for (int c = lane; c < vchunk; c += 32) cp_async16(rowb + 16u * c, sr + 4 * c);
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    // pass 1: scale, mask (only chunks that need it), row maximum; the masked / scaled scores go back to the row
    float m = kNegMax;
    for (int c = lane; c < vchunk; c += 32) {
      float4 v = lds4(rowb + 16u * c);
      const int k0 = c << 2;
      if (a.scale) { v.x = __fmul_rn(v.x, a.mul); v.y = __fmul_rn(v.y, a.mul); v.z = __fmul_rn(v.z, a.mul); v.w = __fmul_rn(v.w, a.mul); }
      if (km || k0 + 3 >= kvis) {
        uint32_t bits = 0xfu;
        if (km) bits = (km[k0 >> 5] >> (k0 & 31)) & 0xfu;
        if (k0 + 3 >= kvis) bits &= (1u << (kvis - k0)) - 1u;
        v.x = (bits & 1u) ? fmaxf(v.x, kNegMax) : kNegMax; v.y = (bits & 2u) ? fmaxf(v.y, kNegMax) : kNegMax;
        v.z = (bits & 4u) ? fmaxf(v.z, kNegMax) : kNegMax; v.w = (bits & 8u) ? fmaxf(v.w, kNegMax) : kNegMax;
      }
      if (a.scale || km || k0 + 3 >= kvis) sts4(rowb + 16u * c, v);
      m = fmaxf(m, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
    }
    m = warp_max(fmaxf(m, kNegMax));
    // a fully masked row (every score finfo.min) is uniform over ALL keys like the reference's: fill the rest of the row and take it along
    int echunk = vchunk;
    if (m == kNegMax) {
      // (same chunk -> lane mapping as the passes below: a lane only ever touches chunks c = lane mod 32 of the row slot)
      for (int c = lane; c < nchunk; c += 32)
        if (c >= vchunk) sts4(rowb + 16u * c, make_float4(kNegMax, kNegMax, kNegMax, kNegMax));
      echunk = nchunk;
    }
    // pass 2: numerators and their sum
    float l = 0.f;
    for (int c = lane; c < echunk; c += 32) {
      float4 v = lds4(rowb + 16u * c);
      v.x = exp_sm<FAST>(v.x, m); v.y = exp_sm<FAST>(v.y, m); v.z = exp_sm<FAST>(v.z, m); v.w = exp_sm<FAST>(v.w, m);
      sts4(rowb + 16u * c, v);
      l = __fadd_rn(l, __fadd_rn(__fadd_rn(v.x, v.y), __fadd_rn(v.z, v.w)));
    }
    const float inv_l = __frcp_rn(warp_add(l));
    // pass 3: normalise, quantise inside the block of 16 (4 lanes), store; whole warps take part in the shuffles
    const int e32 = (echunk + 31) & ~31;
    for (int c = lane; c < e32; c += 32) {
      float4 e = c < echunk ? lds4(rowb + 16u * c) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 p = make_float4(__fmul_rn(e.x, inv_l), __fmul_rn(e.y, inv_l), __fmul_rn(e.z, inv_l), __fmul_rn(e.w, inv_l));
      uint32_t mb = max(max(f2u(p.x), f2u(p.y)), max(f2u(p.z), f2u(p.w)));
      mb = max(mb, __shfl_xor_sync(0xffffffffu, mb, 1));
      mb = max(mb, __shfl_xor_sync(0xffffffffu, mb, 2));
      if (c < nchunk) {
        const float4 q = quant4_probs<KIND>(p, mb, a.f);
        st_stream2(pr + (c << 2), pack_bf16_rn(q.x, q.y), pack_bf16_rn(q.z, q.w));
      }
    }
    for (int c = e32 + lane; c < zchunk; c += 32) st_stream2(pr + (c << 2), 0u, 0u);       // behind the diagonal: exact zeros
    __syncwarp();                                            // the row slot is rewritten by the next row's cp.async
  }
}

bool g_smq_smem_rows = true;          // false: the register-resident v1 kernel (A/B measurement, tests)

template <int KIND, bool FAST>
int launch_smq2(const SmqArgs& a, cudaStream_t st) {
  const int64_t rows = (int64_t)a.batch * a.Sq;
  const int grid = (int)std::min<int64_t>((rows + 7) / 8, (int64_t)num_sms() * 8);
  LaunchScope ls(kKernSoftmaxQuant, st);
  if (g_smq_smem_rows && a.Sk <= 8192) {
    const int wpc = a.Sk <= 2048 ? 8 : (a.Sk <= 4096 ? 4 : 2);
    const size_t smem = (size_t)wpc * a.Sk * 4;
    static PerDevice<size_t> attr_pd;
    size_t& attr = attr_pd.get();
    if (smem > attr) {
      BQ_CUDA_CHECK(cudaFuncSetAttribute(softmax_quant_smem_kernel<KIND, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
      attr = 65536;
    }
    const int g2 = (int)std::min<int64_t>((rows + wpc - 1) / wpc, (int64_t)num_sms() * 8);
    softmax_quant_smem_kernel<KIND, FAST><<<g2, wpc * 32, smem, st>>>(a, wpc);
    return BQ_OK;
  }
  if (a.Sk <= 512) softmax_quant_kernel<KIND, 4, FAST><<<grid, 256, 0, st>>>(a);
  else if (a.Sk <= 1024) softmax_quant_kernel<KIND, 8, FAST><<<grid, 256, 0, st>>>(a);
  else if (a.Sk <= 2048) softmax_quant_kernel<KIND, 16, FAST><<<grid, 256, 0, st>>>(a);
  else softmax_quant_kernel<KIND, 0, FAST><<<grid, 256, 0, st>>>(a);
  return BQ_OK;
}
template <int KIND>
int launch_smq(const SmqArgs& a, cudaStream_t st) {
  return bq_get_attention_precise_exp() ? launch_smq2<KIND, false>(a, st) : launch_smq2<KIND, true>(a, st);
}

// ------------------------------------------------------------------------------------------------ RoPE + quantise / split
struct RopeSplitArgs {
  const float* q;
  const float* k;
  const float* cs;
  const float* sn;
  const int64_t* pos;
  int64_t table_rows;
  __nv_bfloat16* Qq;          // [B][heads][S][d]
  __nv_bfloat16* Kp;          // [B][3][heads][S][d]
  int B, S, heads, d;
  int64_t ldq, ldk;
  FmtParams fq;
};
// one thread per 16 consecutive features of one token of q (blockIdx.y == 0) or k (blockIdx.y == 1); same arithmetic and order as
// rope_quant_q_kernel (quantize.cu): rn(rn(x * cos) + rn(rot * sin))
__global__ void __launch_bounds__(256) rope_split_kernel(RopeSplitArgs a) {
  const int H = a.heads * a.d, bpt = H >> 4, half = a.d >> 1;
  const int64_t nblk = (int64_t)a.B * a.S * bpt;
  const bool is_k = blockIdx.y == 1;
  const float* src = is_k ? a.k : a.q;
  const int64_t ld = is_k ? a.ldk : a.ldq;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nblk; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t tok = i / bpt;
    const int f0 = (int)(i - tok * bpt) << 4;
    const int head = f0 / a.d, e0 = f0 - head * a.d;
    const bool lo = e0 < half;
    const int b = (int)(tok / a.S), s = (int)(tok - (int64_t)b * a.S);
    const int64_t p = a.pos ? min(max(a.pos[tok], (int64_t)0), a.table_rows - 1) : (int64_t)s;
    const float* x = src + tok * ld + f0;
    const float* xp = x + (lo ? half : -half);
    float y[16];
    if (!a.cs) {                                             // no rotation (OPT-style attention): quantise / split the projections as they are
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 xv = *reinterpret_cast<const float4*>(x + 4 * j);
        y[4 * j] = xv.x; y[4 * j + 1] = xv.y; y[4 * j + 2] = xv.z; y[4 * j + 3] = xv.w;
      }
    }
    const float* c = a.cs + p * a.d + e0;
    const float* sn = a.sn + p * a.d + e0;
#pragma unroll
    for (int j = 0; j < 4 && a.cs; ++j) {
      const float4 xv = *reinterpret_cast<const float4*>(x + 4 * j), pv = *reinterpret_cast<const float4*>(xp + 4 * j);
      const float4 cv = __ldg(reinterpret_cast<const float4*>(c + 4 * j)), sv = __ldg(reinterpret_cast<const float4*>(sn + 4 * j));
      const float r0 = lo ? -pv.x : pv.x, r1 = lo ? -pv.y : pv.y, r2 = lo ? -pv.z : pv.z, r3 = lo ? -pv.w : pv.w;
      y[4 * j] = __fadd_rn(__fmul_rn(xv.x, cv.x), __fmul_rn(r0, sv.x));
      y[4 * j + 1] = __fadd_rn(__fmul_rn(xv.y, cv.y), __fmul_rn(r1, sv.y));
      y[4 * j + 2] = __fadd_rn(__fmul_rn(xv.z, cv.z), __fmul_rn(r2, sv.z));
      y[4 * j + 3] = __fadd_rn(__fmul_rn(xv.w, cv.w), __fmul_rn(r3, sv.w));
    }
    if (!is_k) {
      quantize_signed16_rt(y, a.fq);
      uint4* o = reinterpret_cast<uint4*>(a.Qq + (((int64_t)b * a.heads + head) * a.S + s) * a.d + e0);
      o[0] = make_uint4(pack_bf16_rn(y[0], y[1]), pack_bf16_rn(y[2], y[3]), pack_bf16_rn(y[4], y[5]), pack_bf16_rn(y[6], y[7]));
      o[1] = make_uint4(pack_bf16_rn(y[8], y[9]), pack_bf16_rn(y[10], y[11]), pack_bf16_rn(y[12], y[13]), pack_bf16_rn(y[14], y[15]));
    } else {
      // k = k0 + k1 + k2, each a bf16 (error-free: every residual is exact in fp32 and the third plane holds what is left, <= 2^-25 |k|)
      float h1[16], h2[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float h0 = __bfloat162float(__float2bfloat16_rn(y[j]));
        const float r1 = __fsub_rn(y[j], h0);
        h1[j] = __bfloat162float(__float2bfloat16_rn(r1));
        h2[j] = __fsub_rn(r1, h1[j]);
        y[j] = h0;
      }
      const int64_t plane = (int64_t)a.heads * a.S * a.d;
      __nv_bfloat16* o = a.Kp + (int64_t)b * 3 * plane + ((int64_t)head * a.S + s) * a.d + e0;
#pragma unroll
      for (int pl = 0; pl < 3; ++pl) {
        const float* t = pl == 0 ? y : (pl == 1 ? h1 : h2);
        uint4* op = reinterpret_cast<uint4*>(o + pl * plane);
        op[0] = make_uint4(pack_bf16_rn(t[0], t[1]), pack_bf16_rn(t[2], t[3]), pack_bf16_rn(t[4], t[5]), pack_bf16_rn(t[6], t[7]));
        op[1] = make_uint4(pack_bf16_rn(t[8], t[9]), pack_bf16_rn(t[10], t[11]), pack_bf16_rn(t[12], t[13]), pack_bf16_rn(t[14], t[15]));
      }
    }
  }
}

// v fp32 [B][S][heads * d] (token stride ldv) -> planes of v^T: out[b][plane][head][e][s]  (bf16, s contiguous)
// tile: 64 tokens x 32 features through shared memory; a warp writes 128 contiguous bytes (64 tokens) per feature and plane
__global__ void __launch_bounds__(256) split3_transposed_kernel(const float* __restrict__ v, __nv_bfloat16* __restrict__ out, int B, int S,
                                                                 int heads, int d, int64_t ldv) {
  __shared__ float tile[64][33];
  const int H = heads * d;
  const int tiles_s = (S + 63) / 64, tiles_f = H / 32;
  const int64_t ntiles = (int64_t)B * tiles_s * tiles_f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t plane = (int64_t)heads * d * S;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int tf = (int)(t % tiles_f);
    const int64_t r = t / tiles_f;
    const int ts = (int)(r % tiles_s), b = (int)(r / tiles_s);
    const int s0 = ts * 64, f0 = tf * 32;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int s = s0 + warp + i * 8;
      tile[warp + i * 8][lane] = s < S ? v[((int64_t)b * S + s) * ldv + f0 + lane] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int fl = warp + i * 8, f = f0 + fl;
      const int s = s0 + 2 * lane;
      if (s < S) {                                           // S is even: the pair is whole
        const float x0 = tile[2 * lane][fl], x1 = tile[2 * lane + 1][fl];
        const float a0 = __bfloat162float(__float2bfloat16_rn(x0)), a1 = __bfloat162float(__float2bfloat16_rn(x1));
        const float r0 = __fsub_rn(x0, a0), r1 = __fsub_rn(x1, a1);
        const float b0 = __bfloat162float(__float2bfloat16_rn(r0)), b1 = __bfloat162float(__float2bfloat16_rn(r1));
        const float c0 = __fsub_rn(r0, b0), c1 = __fsub_rn(r1, b1);
        __nv_bfloat16* o = out + (int64_t)b * 3 * plane + (int64_t)f * S + s;          // f = head * d + e
        *reinterpret_cast<uint32_t*>(o) = pack_bf16_rn(a0, a1);
        *reinterpret_cast<uint32_t*>(o + plane) = pack_bf16_rn(b0, b1);
        *reinterpret_cast<uint32_t*>(o + 2 * plane) = pack_bf16_rn(c0, c1);
      }
    }
    __syncthreads();
  }
}

}  // namespace
}  // namespace bq

extern "C" {

void bq_set_softmax_smem_rows(int on) { bq::g_smq_smem_rows = on != 0; }

int bq_softmax_quantize(const bq_format* fp, const float* scores, void* P_bf16, int64_t batch, int64_t heads, int64_t Sq, int64_t Sk,
                        int64_t lds, int64_t ss, int64_t ldp, int64_t sp, float score_div, int32_t causal, const uint32_t* key_mask,
                        int64_t key_mask_words, void* stream) {
  using namespace bq;
  if (!fp || batch < 0 || Sq < 0 || Sk < 0 || heads < 1) return BQ_ERR_BAD_ARG;
  if (batch == 0 || Sq == 0 || Sk == 0) return BQ_OK;
  if (!scores || !P_bf16) return BQ_ERR_BAD_ARG;
  if (fp->kind != BQ_KIND_BLOCK_FP && fp->kind != BQ_KIND_BLOCK_MINIFLOAT && fp->kind != BQ_KIND_BLOCK_LOG) return BQ_ERR_UNSUPPORTED;
  if (fp->block_rows != 1 || fp->block_cols != 16) return BQ_ERR_UNSUPPORTED;
  if (Sk % 16) return BQ_ERR_UNSUPPORTED;                       // whole blocks (the reference zero-pads ragged tails: host falls back)
  if (causal && Sq != Sk) return BQ_ERR_UNSUPPORTED;
  if ((batch % heads) || lds < Sk || ldp < Sk || (lds % 4) || (ldp % 4) || ((uintptr_t)scores % 16) || ((uintptr_t)P_bf16 % 8)) return BQ_ERR_BAD_ARG;
  if (batch > 1 && (ss < Sq * lds || sp < Sq * ldp || (ss % 4) || (sp % 4))) return BQ_ERR_BAD_ARG;
  if (key_mask && (key_mask_words * 32 < Sk || ((uintptr_t)key_mask % 4))) return BQ_ERR_BAD_ARG;
  if (batch * Sq > 0x7fffffffffll || Sq > 0x7fffffff || Sk > 0x7fffffff || batch > 0x7fffffff) return BQ_ERR_UNSUPPORTED;
  if (!(score_div > 0.f)) return BQ_ERR_BAD_ARG;
  SmqArgs a;
  memset(&a, 0, sizeof(a));
  int rc = make_params(fp, &a.f);
  if (rc) return rc;
  a.f.fold_zero = 0;
  a.s = scores; a.p = (__nv_bfloat16*)P_bf16; a.lds = lds; a.ss = ss; a.ldp = ldp; a.sp = sp;
  a.batch = (int)batch; a.Sq = (int)Sq; a.Sk = (int)Sk; a.heads = (int)heads;
  if (causal < 0 || causal > 2) return BQ_ERR_BAD_ARG;
  a.mul = 1.0f / score_div; a.scale = score_div != 1.0f; a.causal = causal;
  a.kmask = key_mask; a.kwords = (int)key_mask_words;
  cudaStream_t st = (cudaStream_t)stream;
  if (fp->kind == BQ_KIND_BLOCK_FP) rc = launch_smq<kBlockFP>(a, st);
  else if (fp->kind == BQ_KIND_BLOCK_MINIFLOAT) rc = launch_smq<kBlockMinifloat>(a, st);
  else rc = launch_smq<kBlockLog>(a, st);
  BQ_CUDA_CHECK(cudaGetLastError());
  return rc;
}

int bq_rope_quantize_split(const float* q, const float* k, const float* cos_table, const float* sin_table, const int64_t* position_ids,
                           int64_t table_rows, int64_t B, int64_t S, int32_t heads, int32_t head_dim, int64_t ldq, int64_t ldk,
                           const bq_format* fq, void* Qq_bf16, void* K_planes_bf16, void* stream) {
  using namespace bq;
  if (B < 0 || S < 0 || heads <= 0 || head_dim <= 0 || !fq) return BQ_ERR_BAD_ARG;
  if (B == 0 || S == 0) return BQ_OK;
  if (!q || !k || !Qq_bf16 || !K_planes_bf16 || ((cos_table == nullptr) != (sin_table == nullptr))) return BQ_ERR_BAD_ARG;   // both tables NULL: no rotation
  if (head_dim % 32) return BQ_ERR_UNSUPPORTED;
  if (cos_table && (table_rows < 1 || (!position_ids && table_rows < S))) return BQ_ERR_BAD_ARG;
  const int64_t H = (int64_t)heads * head_dim;
  if (ldq < H || ldk < H || (ldq % 4) || (ldk % 4) || ((uintptr_t)q % 16) || ((uintptr_t)k % 16) || ((uintptr_t)cos_table % 16) ||
      ((uintptr_t)sin_table % 16) || ((uintptr_t)Qq_bf16 % 16) || ((uintptr_t)K_planes_bf16 % 16))
    return BQ_ERR_BAD_ARG;
  if (B * S * H > 0x7fffffffffffll || H > 0x7fffffff) return BQ_ERR_UNSUPPORTED;
  if (fq->kind != BQ_KIND_BLOCK_FP && fq->kind != BQ_KIND_BLOCK_MINIFLOAT && fq->kind != BQ_KIND_BLOCK_LOG) return BQ_ERR_UNSUPPORTED;
  if (fq->block_rows != 1 || fq->block_cols != 16) return BQ_ERR_UNSUPPORTED;
  RopeSplitArgs a;
  memset(&a, 0, sizeof(a));
  int rc = make_params(fq, &a.fq);
  if (rc) return rc;
  a.fq.fold_zero = 0;
  a.q = q; a.k = k; a.cs = cos_table; a.sn = sin_table; a.pos = position_ids; a.table_rows = table_rows;
  a.Qq = (__nv_bfloat16*)Qq_bf16; a.Kp = (__nv_bfloat16*)K_planes_bf16;
  a.B = (int)B; a.S = (int)S; a.heads = heads; a.d = head_dim; a.ldq = ldq; a.ldk = ldk;
  const int64_t n = B * S * (H / 16);
  const int gx = (int)std::min<int64_t>((n + 255) / 256, (int64_t)num_sms() * 16);
  cudaStream_t st = (cudaStream_t)stream;
  {
    LaunchScope ls(kKernRopeSplit, st);
    rope_split_kernel<<<dim3(gx, 2, 1), 256, 0, st>>>(a);
  }
  BQ_CUDA_CHECK(cudaGetLastError());
  return BQ_OK;
}

int bq_split3_bf16_transposed(const float* v, void* planes_bf16, int64_t B, int64_t S, int32_t heads, int32_t head_dim, int64_t ldv,
                              void* stream) {
  using namespace bq;
  if (B < 0 || S < 0 || heads <= 0 || head_dim <= 0) return BQ_ERR_BAD_ARG;
  if (B == 0 || S == 0) return BQ_OK;
  if (!v || !planes_bf16) return BQ_ERR_BAD_ARG;
  const int64_t H = (int64_t)heads * head_dim;
  if ((H % 32) || (S % 2)) return BQ_ERR_UNSUPPORTED;
  if (ldv < H || ((uintptr_t)v % 4) || ((uintptr_t)planes_bf16 % 4)) return BQ_ERR_BAD_ARG;
  if (B > 0x7fffffff || S > 0x7fffffff || H > 0x7fffffff) return BQ_ERR_UNSUPPORTED;
  const int64_t ntiles = B * ((S + 63) / 64) * (H / 32);
  const int grid = (int)std::min<int64_t>(ntiles, (int64_t)num_sms() * 8);
  cudaStream_t st = (cudaStream_t)stream;
  {
    LaunchScope ls(kKernSplit3T, st);
    split3_transposed_kernel<<<grid, 256, 0, st>>>(v, (__nv_bfloat16*)planes_bf16, (int)B, (int)S, heads, head_dim, ldv);
  }
  BQ_CUDA_CHECK(cudaGetLastError());
  return BQ_OK;
}

}  // extern "C"
