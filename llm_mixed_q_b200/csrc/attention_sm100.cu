// attention_sm100.cu — fused causal quantized attention for sm_100a.
//
// Replaces, for one decoder layer, the reference's chain (models/opt_quantized/modeling_opt.py:246-312,
// models/llama_quantized/modeling_llama.py:309-344):
//     scores = bmm(Qx(q), Qy(k^T))            fp32 [B*h, S, S]  written to HBM
//     scores = max(scores + causal_mask, finfo.min) ; probs = softmax(scores)       3 more HBM round trips
//     out    = bmm(Qx(probs), Qy(v))          probs quantised in 1x16 blocks along the key dim
// (and, optionally, the x-quantizer of the following out_proj / o_proj Linear, quantized_modules/linear.py:63-71)
// with ONE kernel in which scores and probabilities never leave the SM:
//     S = Q K^T on tcgen05 (bf16 operands = exact block-quantised q / k, fp32 accumulate in TMEM),
//     sweep 0: running row max / row sum of exp(s - max) (online rescaling), sweep 1: p = exp(s - max) * (1/sum)
//     (S is recomputed in the second sweep — the tensor pipe is far from being the bottleneck),
//     P quantised in registers (one thread owns a query row, so a 1x16 block is 16 consecutive registers),
//     written as a swizzled bf16 K-major smem tile and multiplied with V (MN-major operand) into a TMEM
//     accumulator; the epilogue optionally block-quantises O for the next Linear and emits bf16.
// Blocks of P need FINAL probabilities, which is why this is a two-sweep rather than an online-softmax
// (flash) schedule.  Key tiles above the diagonal are skipped: their probabilities are exactly 0 in the
// reference (exp(finfo.min - max) == 0) and quantise to 0.
//
// The kernel is ALU-issue bound (two exponentials and one block quantisation per score), so the inner loops are
// written for instruction count: the statistics sweep uses ex2.approx on (s - max) * log2(e) (only the row SUM
// depends on it), the final sweep uses libdevice expf like torch's softmax, rounding to the mantissa grid is the
// exact magic-constant add/sub (== rintf for 0 <= t < 2^22), bf16 packing is a byte permute (quantised values have
// <= 8 significant bits), and the causal predicate only exists in the code path of the diagonal tile.
//
// Operands are produced by bq_quantize or by the quantising GEMM epilogues: Qq, Kq, Vq are bf16 [B, S, h, d]
// (token stride given), Kq blocked along S (k^T's last dim), Vq along d.  d in {64, 128}.
//
// Numerics vs the reference's torch softmax: same exp(x - max) with libdevice expf for the numerators; the row sum
// is accumulated in a different order from 2-ulp exponentials and p uses one multiplication by the correctly rounded
// reciprocal instead of a division.  An ulp-level difference only matters when a probability sits on a
// rounding boundary of the block format (one quantisation step there); tests/test_gpu_consumers.py states the
// tolerance.  Pass-through probabilities (p <= 1e-8, returned unquantised by the reference) are truncated to bf16.
//
// Warp roles (640 threads, 1 CTA/SM, persistent):
//   warp 0      TMA producer            warp 1   MMA issuer           warp 2   TMEM allocator
//   warps 4-19  softmax/quantise: warp w owns TMEM lane quarter (w % 4) = 32 query rows and key columns
//               [32*cq, 32*cq+32) of every 128-key tile, cq = (w - 4) / 4
// Work item = (b, h, pair p): query tile T-1-p followed by query tile p (equal cost for every item; the items of
// one head are adjacent in the round-robin order, so ~8 neighbouring CTAs share that head's K / V through L2).
#include "bq_internal.h"
#include "bq_blockops.cuh"
#include "sm100_ptx.cuh"

#include <cuda_bf16.h>
#include <math.h>

namespace bq {

constexpr int kAtBM = 128;      // query rows per work item (UMMA M)
constexpr int kAtBN = 128;      // keys per tile
constexpr int kAtThreads = 640;                    // 4 control warps + 16 softmax warps.  (Dropping the single-pipeline kernel's idle 20th warp
                                                   // does not raise the 96-register cap: the file is 16 K registers per SM sub-partition and one
                                                   // of them still hosts 5 warps — ptxas -v, 608 threads: 96 registers, same spills.)
constexpr int kAt1Threads = kAtThreads;
constexpr int kAt1FirstSoftmaxWarp = 4;
constexpr int kSoftmaxWarps = 16;
constexpr int kSubTile = 128 * 64 * 2;             // one 128-row x 128-byte swizzle-128B sub-tile = 16 KB
constexpr uint32_t kAtTmemCols = 512;              // S: 2 x 128, O: up to 128  -> next power of two
constexpr uint32_t kTmemO = 256;
constexpr float kL2E = 1.4426950408889634f;
constexpr float kMagic = 12582912.0f;              // 1.5 * 2^23: (t + kMagic) - kMagic == rintf(t) for 0 <= t < 2^22

template <int D> struct AtCfg {
  static constexpr int kQBufs = (D == 64) ? 2 : 1;
  static constexpr int kKStages = (D == 64) ? 3 : 2;
  static constexpr int kVStages = (D == 64) ? 2 : 1;
  static constexpr int kTile = (D / 64) * kSubTile;               // Q / K / V tile bytes
  static constexpr int kSmemQ = 0;
  static constexpr int kSmemK = kSmemQ + kQBufs * kTile;
  static constexpr int kSmemV = kSmemK + kKStages * kTile;
  static constexpr int kSmemP = kSmemV + kVStages * kTile;        // 2 buffers x 2 sub-tiles (64 keys each)
  static constexpr int kSmemX = kSmemP + 4 * kSubTile;            // row-stat exchange: 2 x (m, l) x 4 column quarters x 128 rows
  static constexpr int kSmemBar = kSmemX + 2 * 2 * 4 * 128 * 4;            // double-buffered
  static constexpr int kNumBars = 2 * kQBufs + 2 * kKStages + 2 * kVStages + 4 + 4 + 2;
  static constexpr int kSmemBytes = kSmemBar + kNumBars * 8 + 16 + 1024;
};

struct AttnArgs {
  void* out;            // fp32 [B,S,H,d] (out_mode 0) or bf16 (out_mode 1)
  int B, H, S;
  int64_t ldo;          // token stride of out (elements)
  int q_tiles;          // ceil(S / 128)
  int scale_mode;       // 0: none; 1: multiply by score_mul = 1.0f / score_div — what torch-CUDA does for `tensor / python_float`
                        // (BinaryDivTrueKernel: a * (1/b) for a CPU-scalar divisor); identical to a division for powers of two
  float score_mul;
  int out_mode;         // 0: fp32 unquantised; 1: bf16, block-quantised with `po` (blocks of 16 along d)
  FmtParams p;          // format of P (data_in of bmm_1 / matmul_1)
  FmtParams po;         // format of the output (data_in of the following Linear), out_mode 1
  int causal;           // 1: keys above the diagonal are masked (decoder); 0: bidirectional (BERT, modeling_bert.py:366-435)
  const uint32_t* kmask;   // key-validity bitmap [B][kmask_words], bit i of word w = key 32 * w + i takes part; nullptr: all keys < S
  int kmask_words;      // words per batch row (covers whole 128-key tiles; keys >= S are 0)
};

// MN-major SWIZZLE_128B operand: rows of 128 bytes run along MN (64 bf16), 8 such rows (8 K indices) per
// 1024-byte atom.  SBO = distance between 8-K groups; LBO = distance between 64-element MN chunks.
__device__ __forceinline__ uint64_t smem_desc_sw128_mnmajor(uint32_t smem_addr, uint32_t lbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
         ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t idesc_bf16_f32_bmn(int M, int N) {   // B operand MN-major (bit 16)
  return ptx::idesc_bf16_f32(M, N) | (1u << 16);
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// one FMNMX3 each; inline PTX (sm_100 three-input max / min) so that the compiler cannot re-associate the trees below into chains
__device__ __forceinline__ float max3f(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float min3f(float a, float b, float c) {
  float d;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t bf16x2_fma(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint32_t bf16x2_min(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("min.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
// {lo, hi} fp32 -> packed bf16x2, round to nearest even (one F2FP)
__device__ __forceinline__ uint32_t cvt_bf16x2_rn(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
// Build-time experiment, OFF: rounding to the mantissa grid by conversion (see quantize_probs16).  Measured 0.5705 -> 0.5575 ms at
// B8 h32 S2048, but the sum is rounded at 2^-16 before the integer rounding: about one probability in 10^5 lands on the other side
// of a half-integer, where the magic-constant path (one rounding at t's own ulp, like the reference) showed none — on a small W4
// model the fused forward stopped being bit-identical to the op-by-op one (tests/test_gpu_fused_glue.py).  Parity first: not shipped.
#ifndef BQ_PROBS_CVT_ROUND
#define BQ_PROBS_CVT_ROUND 0
#endif
constexpr bool kProbsCvtRound = BQ_PROBS_CVT_ROUND != 0;
// kMagicB = 1.5 * 2^23 + 0x4300: (t + kMagicB) - kMagicB == rintf(t) like kMagic, and the LOW 16 bits of the sum are
// 0x4300 + q — the bf16 encoding of 128 + q (q <= 128).  Two such halves byte-permuted into one word are a bf16x2 pair on
// which the clamp to qmax and the de-quantisation q * 2^(E-m) = fma(128 + q, step, -128 * step) run packed (every value
// exact in bf16: q has <= 7 bits).
constexpr float kMagicB = 12600064.0f;

// Quantise 16 consecutive NON-NEGATIVE values (one reference block of probabilities) and pack them as bf16.
// SCALED: v holds the UN-normalised exponentials e_i and the probabilities are p_i = rn(e_i * inv_l); the common path never
// forms p_i: block max / min scale once (rounding is monotonic) and the normalisation rides on the quantiser's scale,
// t_i = fma(e_i, inv_l * 2^(m-E), c0) — one rounding instead of rn(rn(e_i * inv_l) * 2^(m-E) + c0), i.e. <= 1 ulp of t_i, the
// same class of deviation as the approximate exponential itself.  !SCALED: v holds p_i, inv_l unused.
// Probabilities are <= 1 mathematically; an approximate exponential may overshoot by an ulp, so the BLOCK MAX is clamped to
// 1.0 before the shared exponent is derived (an overshooting element then saturates at qmax exactly like the reference's 1.0).
// Cold tail of quantize_probs16 (a block with a pass-through element, a block whose maximum sits on a log2 cliff or has an extreme
// exponent, block_minifloat): OUT OF LINE, operands through local memory, so that the hot loop body stays a few hundred contiguous
// instructions — with two pipelines executing different loops at the same time the inlined cold code (16 KB per copy) pushed the
// loops out of the 32 KB instruction cache (ncu: no_instruction stalls 1.8 per issue against 0.2 for one pipeline).
template <int KIND, bool SCALED>
__device__ __forceinline__ void quantize_probs16_general(float* v, float inv_l, uint32_t mbits, const FastState& fs, const FmtParams& p,
                                                         uint32_t* w) {
  if (SCALED) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __fmul_rn(v[i], inv_l);
  }
  if (fs.ok) {
    if (KIND == kBlockFP) {
      const float c0 = __fmul_rn(1e-9f, fs.f0);
      const float hi = __fadd_rn(kMagic, p.qmax);       // clamp bound in the magic-shifted domain (exact integer)
      const float c1 = -__fmul_rn(kMagic, fs.f1);       // exact: f1 is a power of two
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float t = __fmaf_rn(v[i], fs.f0, c0);     // == (v + 1e-9f) * 2^(m-E): scaling by 2^k commutes with rounding
        const float tm = fminf(__fadd_rn(t, kMagic), hi);   // kMagic + min(rint(t), qmax)
        const float y = __fmaf_rn(tm, fs.f1, c1);       // (tm - kMagic) * 2^(E-m): every term exact, so is the fma
        v[i] = (v[i] <= 1e-8f) ? v[i] : y;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = quant_elem_fast<KIND>(v[i], fs, p);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = quant_literal_1<KIND>(v[i], mbits, p);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) w[i] = pack_bf16_trunc(v[2 * i], v[2 * i + 1]);
}
template <int KIND, bool SCALED>
__device__ __noinline__ void quantize_probs16_cold(float* v, float inv_l, uint32_t mbits, FastState fs, FmtParams p, uint32_t* w) {
  quantize_probs16_general<KIND, SCALED>(v, inv_l, mbits, fs, p, w);
}

// Quantise 16 consecutive NON-NEGATIVE values (one reference block of probabilities) and pack them as bf16.
// SCALED: v holds the UN-normalised exponentials e_i and the probabilities are p_i = rn(e_i * inv_l); the common path never
// forms p_i: block max / min scale once (rounding is monotonic) and the normalisation rides on the quantiser's scale,
// t_i = fma(e_i, inv_l * 2^(m-E), c0) — one rounding instead of rn(rn(e_i * inv_l) * 2^(m-E) + c0), i.e. <= 1 ulp of t_i, the
// same class of deviation as the approximate exponential itself.  !SCALED: v holds p_i, inv_l unused.
// Probabilities are <= 1 mathematically; an approximate exponential may overshoot by an ulp, so the BLOCK MAX is clamped to
// 1.0 before the shared exponent is derived (an overshooting element then saturates at qmax exactly like the reference's 1.0).
template <int KIND, bool SCALED>
__device__ __forceinline__ void quantize_probs16(float (&v)[16], float inv_l, const FmtParams& p, uint32_t (&w)[8]) {
  // block max / min as trees of 3-input operations (see max32_tree)
  float mx = fmaxf(max3f(max3f(v[0], v[1], v[2]), max3f(v[3], v[4], v[5]), max3f(v[6], v[7], v[8])),
                   max3f(max3f(v[9], v[10], v[11]), max3f(v[12], v[13], v[14]), v[15]));
  if (SCALED) mx = __fmul_rn(mx, inv_l);
  if (mx == 0.f) {                                      // all-zero block -> zeros (pass-through)
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = 0u;
    return;
  }
  mx = fminf(mx, 1.0f);
  const uint32_t mbits = f2u(mx);
  const FastState fs = fast_state<KIND>(mbits, p);
  if (KIND == kBlockFP && fs.ok && p.mbits <= 7) {
    float mn = fminf(min3f(min3f(v[0], v[1], v[2]), min3f(v[3], v[4], v[5]), min3f(v[6], v[7], v[8])),
                     min3f(min3f(v[9], v[10], v[11]), min3f(v[12], v[13], v[14]), v[15]));
    if (SCALED) mn = __fmul_rn(mn, inv_l);
    if (mn > 1e-8f) {
      // no pass-through element in this block (the common case): rounding in fp32, clamp + de-quantisation packed in bf16
      const float c0 = __fmul_rn(1e-9f, fs.f0);         // exact: f0 is a power of two
      const float g0 = SCALED ? __fmul_rn(inv_l, fs.f0) : fs.f0;      // exact (power-of-two factor, |E| <= 100)
      const uint32_t step2 = pack_bf16_trunc(fs.f1, fs.f1);
      const float nb = -__fmul_rn(128.0f, fs.f1);
      const uint32_t base2 = pack_bf16_trunc(nb, nb);
      const float top = __fadd_rn(128.0f, p.qmax);
      const uint32_t top2 = pack_bf16_trunc(top, top);
      if (SCALED && kProbsCvtRound) {
        // (experiment, compiled out by default — see BQ_PROBS_CVT_ROUND) 128 + (p + 1e-9f) * 2^(m-E) in ONE fused rounding, then
        // cvt.rn.bf16x2.f32 — bf16 has an ulp of 1 in [128, 256), so the conversion is the round-to-nearest-even to the mantissa grid
        // and packs the pair: 2 FFMA + 1 F2FP instead of 2 FFMA + 2 FADD + 1 PRMT.
        const float c128 = __fadd_rn(c0, 128.0f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t q2 = bf16x2_min(cvt_bf16x2_rn(__fmaf_rn(v[2 * i], g0, c128), __fmaf_rn(v[2 * i + 1], g0, c128)), top2);
          w[i] = bf16x2_fma(q2, step2, base2);
        }
        return;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float t0 = __fadd_rn(__fmaf_rn(v[2 * i], g0, c0), kMagicB);            // kMagicB + rint((p + 1e-9f) * 2^(m-E))
        const float t1 = __fadd_rn(__fmaf_rn(v[2 * i + 1], g0, c0), kMagicB);
        const uint32_t q2 = bf16x2_min(__byte_perm(f2u(t0), f2u(t1), 0x5410), top2);   // {128 + q0, 128 + q1}, clamped to 128 + qmax
        w[i] = bf16x2_fma(q2, step2, base2);
      }
      return;
    }
  }
  if (KIND == kBlockFP && fs.ok) {
    // Blocks that hold pass-through probabilities (p <= 1e-8, which the reference returns unquantised: block_fp.py's isclose(x, 0)
    // blend) — the common case of a PEAKED softmax, i.e. of trained checkpoints.  Two in-line tiers with the arithmetic of the
    // out-of-line general path (bit-identical to it), so that such rows no longer pay a call and a round trip through local
    // memory per block (1.08 ms against 0.578 ms at B8 h32 S2048 when most blocks were of this kind):
    if (mx <= 1e-8f) {
      // (a) the whole block is below the threshold: every element passes through, truncated to bf16
#pragma unroll
      for (int i = 0; i < 8; ++i)
        w[i] = SCALED ? pack_bf16_trunc(__fmul_rn(v[2 * i], inv_l), __fmul_rn(v[2 * i + 1], inv_l)) : pack_bf16_trunc(v[2 * i], v[2 * i + 1]);
      return;
    }
    // (b) mixed block: quantise in fp32 and select per element
    const float c0 = __fmul_rn(1e-9f, fs.f0);
    const float hi = __fadd_rn(kMagic, p.qmax);
    const float c1 = -__fmul_rn(kMagic, fs.f1);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float pi = SCALED ? __fmul_rn(v[i], inv_l) : v[i];
      const float tm = fminf(__fadd_rn(__fmaf_rn(pi, fs.f0, c0), kMagic), hi);
      const float y = __fmaf_rn(tm, fs.f1, c1);
      v[i] = (pi <= 1e-8f) ? pi : y;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = pack_bf16_trunc(v[2 * i], v[2 * i + 1]);
    return;
  }
  if (KIND == kBlockMinifloat) {
    // block_minifloat has no packed path above: this IS its hot path — keep it in registers (out of line it cost the Llama-7B W4A4
    // attention 0.37 -> 0.58 ms per layer: every block went through local memory and a call)
    // The per-element log2 shortcut runs UNCHECKED here (NC = true): each element min-accumulates its distance from the log2 cliff and
    // the block is tested once — the per-element zone test + branch of the checked variant kept the 16 element chains from
    // interleaving.  A block with an element inside the zone (16 * 2^-14 of the blocks) is redone by the checked out-of-line path; v
    // is left untouched for it.
    if (fs.ok) {
      uint32_t zacc = 0xffffffffu;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float a = quant_elem_fast_impl<KIND, true>(SCALED ? __fmul_rn(v[2 * i], inv_l) : v[2 * i], fs, p, zacc);
        const float b = quant_elem_fast_impl<KIND, true>(SCALED ? __fmul_rn(v[2 * i + 1], inv_l) : v[2 * i + 1], fs, p, zacc);
        w[i] = pack_bf16_trunc(a, b);
      }
      if (!zone_hit<KIND>(zacc)) return;
    }
  }
  float vc[16];
  uint32_t wc[8];
#pragma unroll
  for (int i = 0; i < 16; ++i) vc[i] = v[i];
  quantize_probs16_cold<KIND, SCALED>(vc, inv_l, mbits, fs, p, wc);
#pragma unroll
  for (int i = 0; i < 8; ++i) w[i] = wc[i];
}

// statistics of 32 scores: running max m (with mL = rn(m * log2 e)) and running sum l of 2^(s * log2 e - mL).
// l is kept in "mL units": the exact exp(s - m) differs from the accumulated term by the factor 2^-(m * log2 e - mL),
// which is constant per row and applied once in the merge (stat_fixup) — the per-element work is FFMA + EX2 + FADD.
constexpr float kL2ELo = 1.925963033500011e-08f;     // log2(e) - (float)log2(e)
// Scores behind the causal diagonal are pre-set to -inf by the caller (mask_scores32): exp2(-inf) == 0 and max ignores them, so one
// copy of this code serves full and diagonal slices.
__device__ __forceinline__ void mask_scores32(uint32_t (&r)[32], int nvalid) {
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = (i < nvalid) ? r[i] : 0xff800000u;
}
// maximum of 32 values as a TREE of 3-input maxima (depth 4): the compiler's left-to-right chain of 16 dependent FMNMX3 kept a warp
// from issuing anything else for ~80 cycles per slice — with every softmax warp of a pipeline in the same phase that showed up as
// fixed-latency "wait" stalls (ncu: 1.9 warps per issue slot)
__device__ __forceinline__ float max32_tree(const uint32_t (&r)[32]) {
  float t[11];
#pragma unroll
  for (int i = 0; i < 10; ++i) t[i] = max3f(u2f(r[3 * i]), u2f(r[3 * i + 1]), u2f(r[3 * i + 2]));
  t[10] = fmaxf(u2f(r[30]), u2f(r[31]));
  const float u0 = max3f(t[0], t[1], t[2]), u1 = max3f(t[3], t[4], t[5]), u2 = max3f(t[6], t[7], t[8]), u3 = fmaxf(t[9], t[10]);
  return fmaxf(fmaxf(u0, u1), fmaxf(u2, u3));
}
// key-padding mask: bit i of mw clear -> key i of the slice does not take part (the reference adds finfo.min to its score,
// opt_quantized/modeling_opt.py:520-548, bert_quantized/modeling_bert.py:366-370: exp(finfo.min - max) == 0 exactly like -inf here)
__device__ __forceinline__ void mask_scores32_bits(uint32_t (&r)[32], uint32_t mw) {
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = ((mw >> i) & 1u) ? r[i] : 0xff800000u;
}
template <bool MASKED>
__device__ __forceinline__ void stats32(const uint32_t (&r)[32], float& m, float& mL, float& l) {
  const float tmax = max32_tree(r);
  // a row whose 32 keys of this slice are all masked (key-padding holes) before it has met a valid key: nothing to add, and
  // (-inf) * log2 e - (-inf) below would be NaN
  if (MASKED && tmax == -INFINITY) return;
  if (tmax > m) {                                        // online rescale of the running sum
    const float mLn = __fmul_rn(tmax, kL2E);
    l = __fmul_rn(l, ex2_fast(__fsub_rn(mL, mLn)));      // first time: mL = -inf -> factor 0 (l is 0 anyway)
    m = tmax;
    mL = mLn;
  }
  const float nmL = -mL;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    a0 = __fadd_rn(a0, ex2_fast(__fmaf_rn(u2f(r[i]), kL2E, nmL)));
    a1 = __fadd_rn(a1, ex2_fast(__fmaf_rn(u2f(r[i + 1]), kL2E, nmL)));
    a2 = __fadd_rn(a2, ex2_fast(__fmaf_rn(u2f(r[i + 2]), kL2E, nmL)));
    a3 = __fadd_rn(a3, ex2_fast(__fmaf_rn(u2f(r[i + 3]), kL2E, nmL)));
  }
  l = __fadd_rn(l, __fadd_rn(__fadd_rn(a0, a1), __fadd_rn(a2, a3)));
}
// l in mL units -> sum of exp(s - m):  multiply by 2^-(m * log2(e) - mL), the product's rounding error recovered by FMA
__device__ __forceinline__ float stat_fixup(float m, float mL, float l) {
  if (l == 0.f) return 0.f;                              // slice without a valid key (m = -inf)
  const float err = __fadd_rn(__fmaf_rn(m, kL2E, -mL), __fmul_rn(m, kL2ELo));
  return __fmul_rn(l, ex2_fast(-err));
}

// final probabilities of 32 scores -> two quantised blocks -> swizzled bf16 rows of the P tile.
// FAST: p = ex2.approx(fma(s, log2 e, -mL)) * inv_l (the product folded into the quantiser's scale) — the exponential of the statistics sweep again (2 instructions
// instead of libdevice expf's 8 + the subtraction); inv_l already carries the row constant 2^-(m * log2 e - mL).  Relative error
// of p <= ~(3 + |s - m| * 1.44) ulp instead of <= ~3 ulp; a probability only changes when it sits that close to a rounding
// boundary.  The largest probability may overshoot 1.0 by an ulp: quantize_probs16 clamps the block max.
template <int KIND, bool FAST>
__device__ __forceinline__ void probs32_regs(const uint32_t (&r)[32], float m, float mL, float inv_l, const FmtParams& p,
                                             uint32_t (&w)[16]) {
  const float nmL = -mL;
#pragma unroll
  for (int blk = 0; blk < 2; ++blk) {
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (FAST) {
        v[i] = ex2_fast(__fmaf_rn(u2f(r[blk * 16 + i]), kL2E, nmL));       // normalised inside quantize_probs16; masked score (-inf) -> 0
      } else {
        const float e = expf(__fsub_rn(u2f(r[blk * 16 + i]), m));           // expf(-inf) == 0 for masked scores
        v[i] = __fmul_rn(e, inv_l);
      }
    }
    uint32_t wb[8];
    quantize_probs16<KIND, FAST>(v, inv_l, p, wb);
#pragma unroll
    for (int i = 0; i < 8; ++i) w[blk * 8 + i] = wb[i];
  }
}
template <int KIND, bool FAST>
__device__ __forceinline__ void probs32(const uint32_t (&r)[32], float m, float mL, float inv_l, const FmtParams& p,
                                        uint32_t prow, int chunk0, int sw) {
  uint32_t w[16];
  probs32_regs<KIND, FAST>(r, m, mL, inv_l, p, w);
#pragma unroll
  for (int blk = 0; blk < 2; ++blk) {
    const int chunk = chunk0 + blk * 2;
    sts_v4(prow + ((chunk ^ sw) << 4), w[blk * 8], w[blk * 8 + 1], w[blk * 8 + 2], w[blk * 8 + 3]);
    sts_v4(prow + (((chunk + 1) ^ sw) << 4), w[blk * 8 + 4], w[blk * 8 + 5], w[blk * 8 + 6], w[blk * 8 + 7]);
  }
}

struct Ring {
  int idx = 0;
  uint32_t phase = 0;
  __device__ __forceinline__ void advance(int n) {
    if (++idx == n) { idx = 0; phase ^= 1; }
  }
};

// MASKED = false: purely causal, no key-padding bitmap — the instance the benchmarked decoder path runs; the mask arithmetic of the
// general instance (bidirectional mode, bitmap loads, per-bit selects) is compiled out of it
template <int KIND, int D, bool FAST, bool MASKED>
__global__ void __launch_bounds__(kAt1Threads, 1)
attention_causal_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmV, AttnArgs g) {
  using Cfg = AtCfg<D>;
  const bool causal_m = MASKED ? (g.causal != 0) : true;
  const uint32_t* const kmask_m = MASKED ? g.kmask : nullptr;
  constexpr int kSub = D / 64;                     // 64-wide sub-tiles along d
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // all shared-memory traffic of this kernel uses 32-bit shared-space addresses (no generic pointers: their window base is
  // re-derived from special registers wherever the compiler rematerialises them — measured 15 instructions per 32-score slice)
  uint32_t sb = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  asm volatile("mov.u32 %0, %0;" : "+r"(sb));            // opaque: keep the base in a register instead of re-deriving it per use
  const uint32_t bar0 = sb + Cfg::kSmemBar;
  // barrier map
  auto q_full = [&](int s) { return bar0 + 8u * s; };
  auto q_empty = [&](int s) { return bar0 + 8u * (Cfg::kQBufs + s); };
  const uint32_t bK = bar0 + 8u * (2 * Cfg::kQBufs);
  auto k_full = [&](int s) { return bK + 8u * s; };
  auto k_empty = [&](int s) { return bK + 8u * (Cfg::kKStages + s); };
  const uint32_t bV = bK + 8u * (2 * Cfg::kKStages);
  auto v_full = [&](int s) { return bV + 8u * s; };
  auto v_empty = [&](int s) { return bV + 8u * (Cfg::kVStages + s); };
  const uint32_t bS = bV + 8u * (2 * Cfg::kVStages);
  auto s_full = [&](int s) { return bS + 8u * s; };
  auto s_empty = [&](int s) { return bS + 8u * (2 + s); };
  auto p_full = [&](int s) { return bS + 8u * (4 + s); };
  auto p_empty = [&](int s) { return bS + 8u * (6 + s); };
  const uint32_t o_full = bS + 8u * 8, o_empty = bS + 8u * 9;
  const uint32_t tmem_slot = bS + 8u * 10;
  const uint32_t xch = sb + Cfg::kSmemX;

  // opaque copy of the thread index: under the 102-register cap the compiler otherwise re-reads SR_TID.X (S2R, ~20 cycles in front
  // of a dependent chain) inside the per-slice loop — 3 % of the kernel's stall samples sat on that chain (ncu source view)
  int tid_reg = (int)threadIdx.x;
  if (D == 64) asm volatile("mov.u32 %0, %0;" : "+r"(tid_reg));      // (d = 128 has no register to spare: pinning spills 350 bytes there)
  const int warp = tid_reg >> 5, lane = tid_reg & 31;
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmQ);
    ptx::prefetch_tmap(&tmK);
    ptx::prefetch_tmap(&tmV);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kQBufs; ++s) { ptx::mbar_init(q_full(s), 1); ptx::mbar_init(q_empty(s), 1); }
    for (int s = 0; s < Cfg::kKStages; ++s) { ptx::mbar_init(k_full(s), 1); ptx::mbar_init(k_empty(s), 1); }
    for (int s = 0; s < Cfg::kVStages; ++s) { ptx::mbar_init(v_full(s), 1); ptx::mbar_init(v_empty(s), 1); }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(s_full(s), 1);
      ptx::mbar_init(s_empty(s), kSoftmaxWarps);
      ptx::mbar_init(p_full(s), kSoftmaxWarps);
      ptx::mbar_init(p_empty(s), 1);
    }
    ptx::mbar_init(o_full, 1);
    ptx::mbar_init(o_empty, kSoftmaxWarps);
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc<kAtTmemCols>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot) : "memory");
  ptx::griddep_launch();                                  // programmatic dependent launch: see gemm_sm100.cu
  ptx::griddep_wait();

  const int T = g.q_tiles;
  const int pairs = causal_m ? (T + 1) >> 1 : T;          // bidirectional: every query tile sees every key tile, no pairing needed
  const int items = g.B * g.H * pairs;
  // item -> (b, h, first query tile, number of query tiles); second query tile = T - 1 - first
  auto decode = [&](int w, int& b, int& h, int& qt_hi, int& nsub) {
    const int bh = w / pairs, pr = w - bh * pairs;
    b = bh / g.H;
    h = bh - b * g.H;
    qt_hi = causal_m ? T - 1 - pr : pr;
    nsub = (causal_m && qt_hi != pr) ? 2 : 1;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      Ring qr, kr, vr;
      for (int w = blockIdx.x; w < items; w += gridDim.x) {
        int b, h, qt_hi, nsub;
        decode(w, b, h, qt_hi, nsub);
        for (int sub = 0; sub < nsub; ++sub) {
          const int qt = sub ? (T - 1 - qt_hi) : qt_hi;
          const int n = causal_m ? qt + 1 : T;
          ptx::mbar_wait(q_empty(qr.idx), qr.phase ^ 1);
          ptx::mbar_expect_tx(q_full(qr.idx), Cfg::kTile);
#pragma unroll
          for (int c = 0; c < kSub; ++c)
            tma_load_4d(sb + Cfg::kSmemQ + qr.idx * Cfg::kTile + c * kSubTile, &tmQ, q_full(qr.idx), c * 64, h, qt * kAtBM, b);
          qr.advance(Cfg::kQBufs);
          for (int sweep = 0; sweep < 2; ++sweep) {
            for (int j = 0; j < n; ++j) {
              ptx::mbar_wait(k_empty(kr.idx), kr.phase ^ 1);
              ptx::mbar_expect_tx(k_full(kr.idx), Cfg::kTile);
#pragma unroll
              for (int c = 0; c < kSub; ++c)
                tma_load_4d(sb + Cfg::kSmemK + kr.idx * Cfg::kTile + c * kSubTile, &tmK, k_full(kr.idx), c * 64, h, j * kAtBN, b);
              kr.advance(Cfg::kKStages);
              if (sweep == 1) {
                ptx::mbar_wait(v_empty(vr.idx), vr.phase ^ 1);
                ptx::mbar_expect_tx(v_full(vr.idx), Cfg::kTile);
#pragma unroll
                for (int c = 0; c < kSub; ++c)
                  tma_load_4d(sb + Cfg::kSmemV + vr.idx * Cfg::kTile + c * kSubTile, &tmV, v_full(vr.idx), c * 64, h, j * kAtBN, b);
                vr.advance(Cfg::kVStages);
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idS = ptx::idesc_bf16_f32(kAtBM, kAtBN);
      constexpr uint32_t idO = idesc_bf16_f32_bmn(kAtBM, D);
      Ring qr, kr, vr, sr, pr;
      uint32_t ophase = 0;
      for (int w = blockIdx.x; w < items; w += gridDim.x) {
        int b, h, qt_hi, nsub;
        decode(w, b, h, qt_hi, nsub);
        for (int sub = 0; sub < nsub; ++sub) {
          const int qt = sub ? (T - 1 - qt_hi) : qt_hi;
          const int n = causal_m ? qt + 1 : T;
          ptx::mbar_wait(q_full(qr.idx), qr.phase);
          const uint32_t qbase = sb + Cfg::kSmemQ + qr.idx * Cfg::kTile;
          auto issue_S = [&]() {
            ptx::mbar_wait(k_full(kr.idx), kr.phase);
            ptx::mbar_wait(s_empty(sr.idx), sr.phase ^ 1);
            ptx::tc_fence_after();
            const uint32_t kbase = sb + Cfg::kSmemK + kr.idx * Cfg::kTile;
#pragma unroll
            for (int k = 0; k < D / 16; ++k) {
              const uint64_t qdesc = ptx::smem_desc_sw128_kmajor(qbase + (k >> 2) * kSubTile) + (uint64_t)(2 * (k & 3));
              const uint64_t kdesc = ptx::smem_desc_sw128_kmajor(kbase + (k >> 2) * kSubTile) + (uint64_t)(2 * (k & 3));
              ptx::umma_bf16(tmem + (uint32_t)(sr.idx * kAtBN), qdesc, kdesc, idS, k != 0);
            }
            ptx::umma_commit(k_empty(kr.idx));
            ptx::umma_commit(s_full(sr.idx));
            kr.advance(Cfg::kKStages);
            sr.advance(2);
          };
          for (int j = 0; j < n; ++j) issue_S();                // statistics sweep
          issue_S();                                            // S(0) of the final sweep
          for (int j = 0; j < n; ++j) {
            if (j + 1 < n) issue_S();                           // keep the softmax warps one tile ahead
            ptx::mbar_wait(v_full(vr.idx), vr.phase);
            ptx::mbar_wait(p_full(pr.idx), pr.phase);
            if (j == 0) { ptx::mbar_wait(o_empty, ophase ^ 1); }
            ptx::tc_fence_after();
            const uint32_t pbase = sb + Cfg::kSmemP + pr.idx * 2 * kSubTile;
            const uint32_t vbase = sb + Cfg::kSmemV + vr.idx * Cfg::kTile;
#pragma unroll
            for (int k = 0; k < kAtBN / 16; ++k) {
              const uint64_t adesc = ptx::smem_desc_sw128_kmajor(pbase + (k >> 2) * kSubTile) + (uint64_t)(2 * (k & 3));
              const uint64_t bdesc = smem_desc_sw128_mnmajor(vbase + k * 16 * 128, kSubTile);
              ptx::umma_bf16(tmem + kTmemO, adesc, bdesc, idO, (j | k) != 0);
            }
            ptx::umma_commit(v_empty(vr.idx));
            ptx::umma_commit(p_empty(pr.idx));
            vr.advance(Cfg::kVStages);
            pr.advance(2);
          }
          ptx::umma_commit(o_full);
          ptx::umma_commit(q_empty(qr.idx));
          qr.advance(Cfg::kQBufs);
          ophase ^= 1;
        }
      }
    }
  } else if (warp >= kAt1FirstSoftmaxWarp) {
    // ------------------------------------------------------------------ softmax / quantise / epilogue
    const int quarter = warp & 3;                 // TMEM lane quarter (a warp may only touch lanes 32 * (warp % 4) ..)
    const int cq = (warp - kAt1FirstSoftmaxWarp) >> 2;      // which 32 key columns of every tile
    const int r_in = quarter * 32 + lane;         // query row inside the tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const int sw = r_in & 7;                      // SWIZZLE_128B: 16-byte chunk index ^= (row & 7)
    const int chunk0 = (cq & 1) * 4;
    Ring sr, pr;
    uint32_t ophase = 0;
    int xbuf = 0;                                 // exchange buffer parity (one per sub-item: no second barrier needed)
    // On the diagonal tile the 32x32 slice (quarter, cq) is fully visible when cq < quarter, fully masked when
    // cq > quarter and needs the per-element predicate only when cq == quarter.
    const bool diag_masked = cq > quarter;
    const bool diag_partial = cq == quarter;
    for (int w = blockIdx.x; w < items; w += gridDim.x) {
      int b, h, qt_hi, nsub;
      decode(w, b, h, qt_hi, nsub);
      for (int sub = 0; sub < nsub; ++sub) {
        const int qt = sub ? (T - 1 - qt_hi) : qt_hi;
        const int n = causal_m ? qt + 1 : T;
        const int row = qt * kAtBM + r_in;
        const int nvalid_d = lane + 1;            // cq == quarter: keys [32*cq, 32*cq + lane] of the diagonal tile
        const uint32_t xm = xch + xbuf * (2 * 4 * 128 * 4);   // [4][128] partial maxima
        const uint32_t xl = xm + 4 * 128 * 4;                 // [4][128] partial sums
        xbuf ^= 1;
        float m = -INFINITY, mL = -INFINITY, l = 0.f, inv_l = 0.f;
        for (int sweep = 0; sweep < 2; ++sweep) {
          for (int j = 0; j < n; ++j) {
            const bool diag = causal_m && (j == n - 1);
            uint32_t mw = 0xffffffffu;                           // key-padding bits of this warp's 32 keys
            if (kmask_m) mw = __ldg(kmask_m + (int64_t)b * g.kmask_words + j * 4 + cq);
            const bool skip = (diag && diag_masked) || mw == 0u;
            ptx::mbar_wait(s_full(sr.idx), sr.phase);
            ptx::tc_fence_after();
            uint32_t r[32];
            if (!skip) {
              ptx::tmem_ld_32x32(tmem + lane_addr + (uint32_t)(sr.idx * kAtBN + cq * 32), r);
              ptx::tmem_ld_wait();
            }
            // the scores are in registers: hand the TMEM buffer back to the MMA warp before the math
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(s_empty(sr.idx));
            sr.advance(2);
            if (!skip && g.scale_mode == 1) {
#pragma unroll
              for (int i = 0; i < 32; ++i) r[i] = f2u(__fmul_rn(u2f(r[i]), g.score_mul));
            }
            if (!skip && diag && diag_partial) mask_scores32(r, nvalid_d);
            if (!skip && mw != 0xffffffffu) mask_scores32_bits(r, mw);
            if (sweep == 0) {
              if (!skip) stats32<MASKED>(r, m, mL, l);
            } else {
              ptx::mbar_wait(p_empty(pr.idx), pr.phase ^ 1);
              const uint32_t prow = sb + Cfg::kSmemP + (pr.idx * 2 + (cq >> 1)) * kSubTile + r_in * 128;
              if (skip) {
#pragma unroll
                for (int c = 0; c < 4; ++c) sts_v4(prow + (((chunk0 + c) ^ sw) << 4), 0u, 0u, 0u, 0u);
              } else {
                probs32<KIND, FAST>(r, m, mL, inv_l, g.p, prow, chunk0, sw);
              }
              ptx::fence_proxy_async_smem();     // generic-proxy smem writes -> visible to the MMA (async proxy)
              __syncwarp();
              if (lane == 0) ptx::mbar_arrive(p_full(pr.idx));
              pr.advance(2);
            }
          }
          if (sweep == 0) {
            // merge the four column quarters of every row:  m = max m_c,  l = sum_c l_c * exp(m_c - m).
            // Only the four warps that share this lane quarter exchange data: one 128-thread named barrier per quarter.
            sts_f32(xm + (cq * 128 + r_in) * 4, m);
            sts_f32(xl + (cq * 128 + r_in) * 4, stat_fixup(m, mL, l));
            named_bar_sync(1 + quarter, 128);
            float mc[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) mc[c] = lds_f32(xm + (c * 128 + r_in) * 4);
            const float mm = fmaxf(fmaxf(mc[0], mc[1]), fmaxf(mc[2], mc[3]));
            float ll = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) ll = __fadd_rn(ll, __fmul_rn(lds_f32(xl + (c * 128 + r_in) * 4), expf(__fsub_rn(mc[c], mm))));
            m = mm;
            // the exact sum contains exp(m - m) = 1, so it is >= 1; rounding in the statistics sweep may land an ulp below,
            // which would make the largest probability exceed 1.0 and jump to the next block exponent
            l = fmaxf(ll, 1.0f);
            inv_l = __frcp_rn(l);
            if (FAST) {
              // final sweep evaluates 2^(s * log2 e - mL) with mL = rn(m * log2 e): fold the row constant 2^-(m * log2 e - mL) into 1 / l
              mL = __fmul_rn(m, kL2E);
              const float err = __fadd_rn(__fmaf_rn(m, kL2E, -mL), __fmul_rn(m, kL2ELo));
              inv_l = __fmul_rn(inv_l, ex2_fast(-err));
            }
          }
        }
        // ---- epilogue: O (128 x D fp32 in TMEM) -> global; this warp owns D/4 of the D columns of its 32 rows
        ptx::mbar_wait(o_full, ophase);
        ophase ^= 1;
        ptx::tc_fence_after();
        constexpr int kOB = D / 64;                // 16-column blocks per thread
        uint32_t ro[kOB][16];
#pragma unroll
        for (int c = 0; c < kOB; ++c) ptx::tmem_ld_32x16(tmem + lane_addr + kTmemO + (uint32_t)(cq * (D / 4) + c * 16), ro[c]);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(o_empty);
        if (row < g.S) {
          const int64_t off = ((int64_t)b * g.S + row) * g.ldo + (int64_t)h * D + cq * (D / 4);
          if (g.out_mode == 0) {
            float* o = reinterpret_cast<float*>(g.out) + off;
#pragma unroll
            for (int c = 0; c < kOB; ++c)
#pragma unroll
              for (int i = 0; i < 16; i += 4)
                *reinterpret_cast<float4*>(o + c * 16 + i) = make_float4(u2f(ro[c][i]), u2f(ro[c][i + 1]), u2f(ro[c][i + 2]), u2f(ro[c][i + 3]));
          } else {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(g.out) + off;
#pragma unroll
            for (int c = 0; c < kOB; ++c) {
              float v[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = u2f(ro[c][i]);
              quantize_signed16_rt(v, g.po);
              uint32_t wv[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) wv[i] = pack_bf16_rn(v[2 * i], v[2 * i + 1]);
              *reinterpret_cast<uint4*>(o + c * 16) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
              *reinterpret_cast<uint4*>(o + c * 16 + 8) = make_uint4(wv[4], wv[5], wv[6], wv[7]);
            }
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<kAtTmemCols>(tmem);
  }
}

// ------------------------------------------------------------------------------------------------
// v4, head_dim 64: TWO independent pipelines ("groups") per CTA and P in tensor memory.
//
// Why: the single-pipeline kernel above is ALU-issue bound but only 59 % issue-active (ncu) — all 16 softmax warps walk through the
// same phases together, so the SM idles whenever they wait together: the hand-over between the two sweeps (row-statistics merge),
// the wait for the last PV MMA before the epilogue, the first S tile of the next item.  Here the 16 softmax warps form two groups
// of 8 that work on DIFFERENT work items with their own TMA lane, MMA lane, barriers, shared-memory ring and TMEM columns: one
// group's bubbles are filled by the other group's math (the two-softmax-warpgroup ping-pong of Blackwell attention kernels).
// P no longer travels registers -> swizzled smem -> fence.proxy.async -> MMA: a softmax warp writes its 32 x 32 quantised
// probabilities (bf16 pairs) straight into tensor memory with tcgen05.st and PV runs with the A operand read FROM TMEM
// (tcgen05.mma [d], [a], b-desc) — no shared-memory store, no generic->async proxy fence on the softmax warps' critical path, and
// 64 KB of shared memory freed for deeper K / V rings.
//
// Key tiles are 64 wide (MMA N = 64) so that one group's accumulators fit 256 TMEM columns:
//   columns (relative to the group's base): S0 [0,64)  S1 [64,128)  O [128,192)  P0 [192,224)  P1 [224,256)
// Warp roles (640 threads): warp 0 / 2 TMA of group 0 / 1 (warp 2 also allocates TMEM), warp 1 / 3 MMA of group 0 / 1,
// warps 4-11 softmax of group 0, 12-19 of group 1; inside a group warp i owns TMEM lane quarter (i & 3) = 32 query rows and key
// columns [32*cq, 32*cq + 32) of every 64-key tile, cq = i >> 2.  Work items are dealt round-robin to the 2 * gridDim.x groups.
// ------------------------------------------------------------------------------------------------
struct At2Cfg {
  static constexpr int kBN = 64;                                     // keys per tile
  static constexpr int kQBufs = 2, kKStages = 4, kVStages = 3;
  static constexpr int kQTile = kAtBM * 64 * 2;                      // 16 KB
  static constexpr int kKVTile = kBN * 64 * 2;                       //  8 KB
  static constexpr int kSmemQ = 0;
  static constexpr int kSmemK = kSmemQ + kQBufs * kQTile;
  static constexpr int kSmemV = kSmemK + kKStages * kKVTile;
  static constexpr int kSmemX = kSmemV + kVStages * kKVTile;         // row-stat exchange: 2 parities x (m, l) x 2 column halves x 128 rows
  static constexpr int kGroupBytes = kSmemX + 2 * 2 * 2 * 128 * 4;
  static constexpr int kBarsPerGroup = 2 * kQBufs + 2 * kKStages + 2 * kVStages + 4 + 4 + 2;
  static constexpr int kSmemBar = 2 * kGroupBytes;
  static constexpr int kSmemBytes = kSmemBar + 2 * kBarsPerGroup * 8 + 16 + 1024;
  static constexpr uint32_t kTmemGroup = 256, kTmemO = 128, kTmemP = 192;
  static constexpr int kGroupWarps = 8;
  static_assert(kGroupBytes % 1024 == 0, "group regions keep the 1024-byte alignment of the swizzle atoms");
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
};

template <int KIND, bool FAST, bool MASKED>
__global__ void __launch_bounds__(kAtThreads, 1)
attention_causal_dual_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                             const __grid_constant__ CUtensorMap tmV, AttnArgs g) {
  using Cfg = At2Cfg;
  const bool causal_m = MASKED ? (g.causal != 0) : true;
  const uint32_t* const kmask_m = MASKED ? g.kmask : nullptr;
  constexpr int D = 64;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint32_t sb0 = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  asm volatile("mov.u32 %0, %0;" : "+r"(sb0));
  int tid_reg = (int)threadIdx.x;
  asm volatile("mov.u32 %0, %0;" : "+r"(tid_reg));
  const int warp = tid_reg >> 5, lane = tid_reg & 31;
  // group of this warp: control warps 0,1 -> 0; 2,3 -> 1; softmax warps 4-11 -> 0; 12-19 -> 1
  const int grp = warp < 4 ? (warp >> 1) : ((warp - 4) >> 3);
  const uint32_t sb = sb0 + (uint32_t)grp * Cfg::kGroupBytes;
  const uint32_t bar0 = sb0 + Cfg::kSmemBar + (uint32_t)grp * (Cfg::kBarsPerGroup * 8);
  auto q_full = [&](int s) { return bar0 + 8u * s; };
  auto q_empty = [&](int s) { return bar0 + 8u * (Cfg::kQBufs + s); };
  const uint32_t bK = bar0 + 8u * (2 * Cfg::kQBufs);
  auto k_full = [&](int s) { return bK + 8u * s; };
  auto k_empty = [&](int s) { return bK + 8u * (Cfg::kKStages + s); };
  const uint32_t bV = bK + 8u * (2 * Cfg::kKStages);
  auto v_full = [&](int s) { return bV + 8u * s; };
  auto v_empty = [&](int s) { return bV + 8u * (Cfg::kVStages + s); };
  const uint32_t bS = bV + 8u * (2 * Cfg::kVStages);
  auto s_full = [&](int s) { return bS + 8u * s; };
  auto s_empty = [&](int s) { return bS + 8u * (2 + s); };
  auto p_full = [&](int s) { return bS + 8u * (4 + s); };
  auto p_empty = [&](int s) { return bS + 8u * (6 + s); };
  const uint32_t o_full = bS + 8u * 8, o_empty = bS + 8u * 9;
  const uint32_t tmem_slot = sb0 + Cfg::kSmemBar + 2 * Cfg::kBarsPerGroup * 8;
  const uint32_t xch = sb + Cfg::kSmemX;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmQ);
    ptx::prefetch_tmap(&tmK);
    ptx::prefetch_tmap(&tmV);
  }
  if ((warp == 1 || warp == 3) && lane == 0) {                    // each MMA lane initialises its own group's barriers
    for (int s = 0; s < Cfg::kQBufs; ++s) { ptx::mbar_init(q_full(s), 1); ptx::mbar_init(q_empty(s), 1); }
    for (int s = 0; s < Cfg::kKStages; ++s) { ptx::mbar_init(k_full(s), 1); ptx::mbar_init(k_empty(s), 1); }
    for (int s = 0; s < Cfg::kVStages; ++s) { ptx::mbar_init(v_full(s), 1); ptx::mbar_init(v_empty(s), 1); }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(s_full(s), 1);
      ptx::mbar_init(s_empty(s), Cfg::kGroupWarps);
      ptx::mbar_init(p_full(s), Cfg::kGroupWarps);
      ptx::mbar_init(p_empty(s), 1);
    }
    ptx::mbar_init(o_full, 1);
    ptx::mbar_init(o_empty, Cfg::kGroupWarps);
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc<kAtTmemCols>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot) : "memory");
  tmem += (uint32_t)grp * Cfg::kTmemGroup;                        // this group's 256 columns
  ptx::griddep_launch();                                          // programmatic dependent launch: see gemm_sm100.cu
  ptx::griddep_wait();

  const int T = g.q_tiles;
  const int pairs = causal_m ? (T + 1) >> 1 : T;
  const int items = g.B * g.H * pairs;
  const int vb = (int)blockIdx.x * 2 + grp, vgrid = (int)gridDim.x * 2;      // virtual worker index: one per group
  const int kt_all = (g.S + Cfg::kBN - 1) / Cfg::kBN;                        // key tiles of a bidirectional row block
  auto decode = [&](int w, int& b, int& h, int& qt_hi, int& nsub) {
    const int bh = w / pairs, pr = w - bh * pairs;
    b = bh / g.H;
    h = bh - b * g.H;
    qt_hi = causal_m ? T - 1 - pr : pr;
    nsub = (causal_m && qt_hi != pr) ? 2 : 1;
  };

  if (warp == 0 || warp == 2) {
    // ------------------------------------------------------------------ TMA producer of this group
    if (lane == 0) {
      Ring qr, kr, vr;
      for (int w = vb; w < items; w += vgrid) {
        int b, h, qt_hi, nsub;
        decode(w, b, h, qt_hi, nsub);
        for (int sub = 0; sub < nsub; ++sub) {
          const int qt = sub ? (T - 1 - qt_hi) : qt_hi;
          const int n = causal_m ? 2 * (qt + 1) : kt_all;                              // 64-key tiles up to and including the diagonal
          ptx::mbar_wait(q_empty(qr.idx), qr.phase ^ 1);
          ptx::mbar_expect_tx(q_full(qr.idx), Cfg::kQTile);
          tma_load_4d(sb + Cfg::kSmemQ + qr.idx * Cfg::kQTile, &tmQ, q_full(qr.idx), 0, h, qt * kAtBM, b);
          qr.advance(Cfg::kQBufs);
          for (int sweep = 0; sweep < 2; ++sweep) {
            for (int j = 0; j < n; ++j) {
              ptx::mbar_wait(k_empty(kr.idx), kr.phase ^ 1);
              ptx::mbar_expect_tx(k_full(kr.idx), Cfg::kKVTile);
              tma_load_4d(sb + Cfg::kSmemK + kr.idx * Cfg::kKVTile, &tmK, k_full(kr.idx), 0, h, j * Cfg::kBN, b);
              kr.advance(Cfg::kKStages);
              if (sweep == 1) {
                ptx::mbar_wait(v_empty(vr.idx), vr.phase ^ 1);
                ptx::mbar_expect_tx(v_full(vr.idx), Cfg::kKVTile);
                tma_load_4d(sb + Cfg::kSmemV + vr.idx * Cfg::kKVTile, &tmV, v_full(vr.idx), 0, h, j * Cfg::kBN, b);
                vr.advance(Cfg::kVStages);
              }
            }
          }
        }
      }
    }
  } else if (warp == 1 || warp == 3) {
    // ------------------------------------------------------------------ MMA issuer of this group
    if (lane == 0) {
      constexpr uint32_t idS = ptx::idesc_bf16_f32(kAtBM, Cfg::kBN);
      constexpr uint32_t idO = idesc_bf16_f32_bmn(kAtBM, D);
      Ring qr, kr, vr, sr, pr;
      uint32_t ophase = 0;
      for (int w = vb; w < items; w += vgrid) {
        int b, h, qt_hi, nsub;
        decode(w, b, h, qt_hi, nsub);
        for (int sub = 0; sub < nsub; ++sub) {
          const int qt = sub ? (T - 1 - qt_hi) : qt_hi;
          const int n = causal_m ? 2 * (qt + 1) : kt_all;
          ptx::mbar_wait(q_full(qr.idx), qr.phase);
          const uint32_t qbase = sb + Cfg::kSmemQ + qr.idx * Cfg::kQTile;
          auto issue_S = [&]() {
            ptx::mbar_wait(k_full(kr.idx), kr.phase);
            ptx::mbar_wait(s_empty(sr.idx), sr.phase ^ 1);
            ptx::tc_fence_after();
            const uint32_t kbase = sb + Cfg::kSmemK + kr.idx * Cfg::kKVTile;
#pragma unroll
            for (int k = 0; k < D / 16; ++k) {
              const uint64_t qdesc = ptx::smem_desc_sw128_kmajor(qbase) + (uint64_t)(2 * k);
              const uint64_t kdesc = ptx::smem_desc_sw128_kmajor(kbase) + (uint64_t)(2 * k);
              ptx::umma_bf16(tmem + (uint32_t)(sr.idx * Cfg::kBN), qdesc, kdesc, idS, k != 0);
            }
            ptx::umma_commit(k_empty(kr.idx));
            ptx::umma_commit(s_full(sr.idx));
            kr.advance(Cfg::kKStages);
            sr.advance(2);
          };
          for (int j = 0; j < n; ++j) issue_S();                // statistics sweep
          issue_S();                                            // S(0) of the final sweep
          for (int j = 0; j < n; ++j) {
            if (j + 1 < n) issue_S();                           // keep the softmax warps one tile ahead
            ptx::mbar_wait(v_full(vr.idx), vr.phase);
            ptx::mbar_wait(p_full(pr.idx), pr.phase);
            if (j == 0) { ptx::mbar_wait(o_empty, ophase ^ 1); }
            ptx::tc_fence_after();
            const uint32_t vbase = sb + Cfg::kSmemV + vr.idx * Cfg::kKVTile;
            const uint32_t ptm = tmem + Cfg::kTmemP + (uint32_t)(pr.idx * (Cfg::kBN / 2));
#pragma unroll
            for (int k = 0; k < Cfg::kBN / 16; ++k) {
              // A: 16 keys = 8 TMEM columns of bf16 pairs; B: 16 key rows (128 bytes each) of the V tile, MN-major
              const uint64_t bdesc = smem_desc_sw128_mnmajor(vbase + k * 16 * 128, Cfg::kKVTile);
              ptx::umma_bf16_ts(tmem + Cfg::kTmemO, ptm + (uint32_t)(8 * k), bdesc, idO, (j | k) != 0);
            }
            ptx::umma_commit(v_empty(vr.idx));
            ptx::umma_commit(p_empty(pr.idx));
            vr.advance(Cfg::kVStages);
            pr.advance(2);
          }
          ptx::umma_commit(o_full);
          ptx::umma_commit(q_empty(qr.idx));
          qr.advance(Cfg::kQBufs);
          ophase ^= 1;
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax / quantise / epilogue
    const int quarter = warp & 3;                 // TMEM lane quarter (== warp % 4)
    const int cq = ((warp - 4) & 7) >> 2;         // which 32 key columns of every 64-key tile
    const int r_in = quarter * 32 + lane;         // query row inside the tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    Ring sr, pr;
    uint32_t ophase = 0;
    int xbuf = 0;
    for (int w = vb; w < items; w += vgrid) {
      int b, h, qt_hi, nsub;
      decode(w, b, h, qt_hi, nsub);
      for (int sub = 0; sub < nsub; ++sub) {
        const int qt = sub ? (T - 1 - qt_hi) : qt_hi;
        const int n = causal_m ? 2 * (qt + 1) : kt_all;
        const int row = qt * kAtBM + r_in;
        const int nvalid_d = lane + 1;
        const uint32_t xm = xch + xbuf * (2 * 2 * 128 * 4);   // [2][128] partial maxima
        const uint32_t xl = xm + 2 * 128 * 4;                 // [2][128] partial sums
        xbuf ^= 1;
        float m = -INFINITY, mL = -INFINITY, l = 0.f, inv_l = 0.f;
        for (int sweep = 0; sweep < 2; ++sweep) {
          for (int j = 0; j < n; ++j) {
            // the last two 64-key tiles straddle the diagonal: 32-key column group c' = 2 * (j - (n - 2)) + cq of the 128 x 128
            // diagonal square is fully visible for c' < quarter, fully masked for c' > quarter, per-element for c' == quarter
            const int cd = (causal_m && j >= n - 2) ? (2 * (j - (n - 2)) + cq) : -1;
            uint32_t mw = 0xffffffffu;                           // key-padding bits of this warp's 32 keys
            if (kmask_m) mw = __ldg(kmask_m + (int64_t)b * g.kmask_words + j * 2 + cq);
            const bool skip = cd > quarter || mw == 0u;
            const bool partial = cd == quarter;
            ptx::mbar_wait(s_full(sr.idx), sr.phase);
            ptx::tc_fence_after();
            uint32_t r[32];
            if (!skip) {
              ptx::tmem_ld_32x32(tmem + lane_addr + (uint32_t)(sr.idx * Cfg::kBN + cq * 32), r);
              ptx::tmem_ld_wait();
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(s_empty(sr.idx));
            sr.advance(2);
            if (!skip && g.scale_mode == 1) {
#pragma unroll
              for (int i = 0; i < 32; ++i) r[i] = f2u(__fmul_rn(u2f(r[i]), g.score_mul));
            }
            if (partial && !skip) mask_scores32(r, nvalid_d);
            if (!skip && mw != 0xffffffffu) mask_scores32_bits(r, mw);
            if (sweep == 0) {
              if (!skip) stats32<MASKED>(r, m, mL, l);
            } else {
              uint32_t wq[16];
              if (skip) {
#pragma unroll
                for (int i = 0; i < 16; ++i) wq[i] = 0u;
              } else {
                probs32_regs<KIND, FAST>(r, m, mL, inv_l, g.p, wq);
              }
              ptx::mbar_wait(p_empty(pr.idx), pr.phase ^ 1);   // the PV MMA that read this P buffer two tiles ago has retired
              ptx::tc_fence_after();
              ptx::tmem_st_32x16(tmem + lane_addr + Cfg::kTmemP + (uint32_t)(pr.idx * (Cfg::kBN / 2) + cq * 16), wq);
              ptx::tmem_st_wait();
              ptx::tc_fence_before();
              __syncwarp();
              if (lane == 0) ptx::mbar_arrive(p_full(pr.idx));
              pr.advance(2);
            }
          }
          if (sweep == 0) {
            // merge the two column halves of every row (the two warps that share this lane quarter inside the group)
            sts_f32(xm + (cq * 128 + r_in) * 4, m);
            sts_f32(xl + (cq * 128 + r_in) * 4, stat_fixup(m, mL, l));
            named_bar_sync(1 + grp * 4 + quarter, 64);
            const float m0 = lds_f32(xm + r_in * 4), m1 = lds_f32(xm + (128 + r_in) * 4);
            const float mm = fmaxf(m0, m1);
            float ll = __fmul_rn(lds_f32(xl + r_in * 4), expf(__fsub_rn(m0, mm)));
            ll = __fadd_rn(ll, __fmul_rn(lds_f32(xl + (128 + r_in) * 4), expf(__fsub_rn(m1, mm))));
            m = mm;
            l = fmaxf(ll, 1.0f);                                 // see the single-pipeline kernel: the exact sum is >= 1
            inv_l = __frcp_rn(l);
            if (FAST) {
              mL = __fmul_rn(m, kL2E);
              const float err = __fadd_rn(__fmaf_rn(m, kL2E, -mL), __fmul_rn(m, kL2ELo));
              inv_l = __fmul_rn(inv_l, ex2_fast(-err));
            }
          }
        }
        // ---- epilogue: O (128 x 64 fp32 in TMEM) -> global; this warp owns 32 of the 64 columns of its 32 rows
        ptx::mbar_wait(o_full, ophase);
        ophase ^= 1;
        ptx::tc_fence_after();
        uint32_t ro[2][16];
#pragma unroll
        for (int c = 0; c < 2; ++c) ptx::tmem_ld_32x16(tmem + lane_addr + Cfg::kTmemO + (uint32_t)(cq * 32 + c * 16), ro[c]);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(o_empty);
        if (row < g.S) {
          const int64_t off = ((int64_t)b * g.S + row) * g.ldo + (int64_t)h * D + cq * 32;
          if (g.out_mode == 0) {
            float* o = reinterpret_cast<float*>(g.out) + off;
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
              for (int i = 0; i < 16; i += 4)
                *reinterpret_cast<float4*>(o + c * 16 + i) = make_float4(u2f(ro[c][i]), u2f(ro[c][i + 1]), u2f(ro[c][i + 2]), u2f(ro[c][i + 3]));
          } else {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(g.out) + off;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              float v[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = u2f(ro[c][i]);
              quantize_signed16_rt(v, g.po);
              uint32_t wv[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) wv[i] = pack_bf16_rn(v[2 * i], v[2 * i + 1]);
              *reinterpret_cast<uint4*>(o + c * 16) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
              *reinterpret_cast<uint4*>(o + c * 16 + 8) = make_uint4(wv[4], wv[5], wv[6], wv[7]);
            }
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<kAtTmemCols>(tmem - (uint32_t)grp * Cfg::kTmemGroup);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
int make_params(const bq_format* f, FmtParams* p);
int make_tmap_bf16_4d(CUtensorMap* tm, const void* base, int64_t d, int64_t S, int64_t H, int64_t B, int64_t ld_tok,
                      int box_rows);

static bool g_attn_precise_exp = false;       // true: libdevice expf for the numerators (bit-identical to torch's exp(x - max))
static bool g_attn_dual = true;               // head_dim 64: the two-pipeline / P-in-TMEM kernel (false: the single-pipeline kernel, A/B)

template <int KIND, bool FAST, bool MASKED>
static int launch_attention_dual_fm(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnArgs& g,
                                    cudaStream_t st) {
  static PerDevice<bool> attr_pd;
  bool& attr = attr_pd.get();
  if (!attr) {
    BQ_CUDA_CHECK(cudaFuncSetAttribute(attention_causal_dual_kernel<KIND, FAST, MASKED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       At2Cfg::kSmemBytes));
    attr = true;
  }
  const int items = g.B * g.H * (g.causal ? (g.q_tiles + 1) / 2 : g.q_tiles);
  const int grid = std::min((items + 1) / 2, num_sms());
  {
    LaunchScope ls(kKernAttention, st);
    BQ_CUDA_CHECK(launch_ex(attention_causal_dual_kernel<KIND, FAST, MASKED>, grid, kAtThreads, At2Cfg::kSmemBytes, st, 1, tq, tk, tv, g));
  }
  BQ_CUDA_CHECK(cudaGetLastError());
  return BQ_OK;
}
template <int KIND, bool FAST>
static int launch_attention_dual_f(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnArgs& g,
                                   cudaStream_t st) {
  return (g.causal && !g.kmask) ? launch_attention_dual_fm<KIND, FAST, false>(tq, tk, tv, g, st)
                                : launch_attention_dual_fm<KIND, FAST, true>(tq, tk, tv, g, st);
}
template <int KIND>
static int launch_attention_dual(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnArgs& g, cudaStream_t st) {
  return g_attn_precise_exp ? launch_attention_dual_f<KIND, false>(tq, tk, tv, g, st) : launch_attention_dual_f<KIND, true>(tq, tk, tv, g, st);
}

template <int KIND, int D, bool FAST, bool MASKED>
static int launch_attention_fm(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnArgs& g,
                               cudaStream_t st) {
  using Cfg = AtCfg<D>;
  static PerDevice<bool> attr_pd;
  bool& attr = attr_pd.get();
  if (!attr) {
    BQ_CUDA_CHECK(cudaFuncSetAttribute(attention_causal_kernel<KIND, D, FAST, MASKED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg::kSmemBytes));
    attr = true;
  }
  const int items = g.B * g.H * (g.causal ? (g.q_tiles + 1) / 2 : g.q_tiles);
  const int grid = std::min(items, num_sms());
  {
    LaunchScope ls(kKernAttention, st);
    BQ_CUDA_CHECK(launch_ex(attention_causal_kernel<KIND, D, FAST, MASKED>, grid, kAt1Threads, Cfg::kSmemBytes, st, 1, tq, tk, tv, g));
  }
  BQ_CUDA_CHECK(cudaGetLastError());
  return BQ_OK;
}
template <int KIND, int D, bool FAST>
static int launch_attention_f(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnArgs& g,
                              cudaStream_t st) {
  return (g.causal && !g.kmask) ? launch_attention_fm<KIND, D, FAST, false>(tq, tk, tv, g, st)
                                : launch_attention_fm<KIND, D, FAST, true>(tq, tk, tv, g, st);
}

template <int KIND, int D>
static int launch_attention(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnArgs& g, cudaStream_t st) {
  return g_attn_precise_exp ? launch_attention_f<KIND, D, false>(tq, tk, tv, g, st) : launch_attention_f<KIND, D, true>(tq, tk, tv, g, st);
}

static int attention_impl(const bq_format* fp, const bq_format* fo, const void* Qq, const void* Kq, const void* Vq, void* out,
                          int64_t B, int64_t H, int64_t S, int64_t d, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo,
                          float score_div, cudaStream_t st, int causal = 1, const uint32_t* key_mask = nullptr, int64_t key_mask_words = 0) {
  if (!fp || B < 0 || H < 0 || S < 0) return BQ_ERR_BAD_ARG;
  if (B == 0 || H == 0 || S == 0) return BQ_OK;
  if (!Qq || !Kq || !Vq || !out) return BQ_ERR_BAD_ARG;
  if (d != 64 && d != 128) return BQ_ERR_UNSUPPORTED;
  if (fp->kind != BQ_KIND_BLOCK_FP && fp->kind != BQ_KIND_BLOCK_MINIFLOAT) return BQ_ERR_UNSUPPORTED;
  if (fp->block_rows != 1 || fp->block_cols != 16) return BQ_ERR_UNSUPPORTED;
  if (fo && ((fo->kind != BQ_KIND_BLOCK_FP && fo->kind != BQ_KIND_BLOCK_MINIFLOAT) || fo->block_rows != 1 || fo->block_cols != 16))
    return BQ_ERR_UNSUPPORTED;
  if ((ldq % 8) || (ldk % 8) || (ldv % 8) || (ldo % (fo ? 8 : 4)) || ((uintptr_t)Qq % 16) || ((uintptr_t)Kq % 16) ||
      ((uintptr_t)Vq % 16) || ((uintptr_t)out % 16))
    return BQ_ERR_BAD_ARG;
  if (!(score_div > 0.f) || !isfinite(score_div)) return BQ_ERR_BAD_ARG;
  if (B * H * ((S + 127) / 128) > 0x7fffffffll || S > 0x7fffffff) return BQ_ERR_UNSUPPORTED;
  AttnArgs g;
  memset(&g, 0, sizeof(g));
  int rc = make_params(fp, &g.p);
  if (rc) return rc;
  g.p.fold_zero = 0;
  if (fo) {
    if ((rc = make_params(fo, &g.po))) return rc;
    g.po.fold_zero = 0;
    g.out_mode = 1;
  }
  g.out = out; g.B = (int)B; g.H = (int)H; g.S = (int)S; g.ldo = ldo;
  g.q_tiles = (int)((S + kAtBM - 1) / kAtBM);
  g.causal = causal ? 1 : 0;
  // every 128-key tile must be covered by the bitmap (keys >= S cleared by the caller); bidirectional attention NEEDS one (nothing
  // else masks the zero-filled keys beyond S), causal attention takes it only for padded batches
  if (key_mask) {
    if (key_mask_words < 4 * (int64_t)g.q_tiles || ((uintptr_t)key_mask % 4)) return BQ_ERR_BAD_ARG;
    g.kmask = key_mask; g.kmask_words = (int)key_mask_words;
  } else if (!causal) {
    return BQ_ERR_BAD_ARG;
  }
  g.score_mul = 1.0f / score_div;
  g.scale_mode = (score_div == 1.0f) ? 0 : 1;
  CUtensorMap tq, tk, tv;
  const bool dual = g_attn_dual && d == 64;
  const int kv_rows = dual ? At2Cfg::kBN : kAtBN;
  if ((rc = make_tmap_bf16_4d(&tq, Qq, d, S, H, B, ldq, kAtBM))) return rc;
  if ((rc = make_tmap_bf16_4d(&tk, Kq, d, S, H, B, ldk, kv_rows))) return rc;
  if ((rc = make_tmap_bf16_4d(&tv, Vq, d, S, H, B, ldv, kv_rows))) return rc;
  const bool bfp = fp->kind == BQ_KIND_BLOCK_FP;
  if (dual) return bfp ? launch_attention_dual<kBlockFP>(tq, tk, tv, g, st) : launch_attention_dual<kBlockMinifloat>(tq, tk, tv, g, st);
  if (d == 64) return bfp ? launch_attention<kBlockFP, 64>(tq, tk, tv, g, st) : launch_attention<kBlockMinifloat, 64>(tq, tk, tv, g, st);
  return bfp ? launch_attention<kBlockFP, 128>(tq, tk, tv, g, st) : launch_attention<kBlockMinifloat, 128>(tq, tk, tv, g, st);
}

}  // namespace bq

extern "C" void bq_set_attention_precise_exp(int on) { bq::g_attn_precise_exp = on != 0; }
extern "C" int bq_get_attention_precise_exp(void) { return bq::g_attn_precise_exp ? 1 : 0; }
extern "C" void bq_set_attention_dual_pipeline(int on) { bq::g_attn_dual = on != 0; }
extern "C" int bq_get_attention_dual_pipeline(void) { return bq::g_attn_dual ? 1 : 0; }

extern "C" int bq_attention_causal(const bq_format* fp, const void* Qq, const void* Kq, const void* Vq, float* out,
                                   int64_t B, int64_t H, int64_t S, int64_t d, int64_t ldq, int64_t ldk, int64_t ldv,
                                   int64_t ldo, float score_div, void* stream) {
  return bq::attention_impl(fp, nullptr, Qq, Kq, Vq, out, B, H, S, d, ldq, ldk, ldv, ldo, score_div, (cudaStream_t)stream);
}

extern "C" int bq_attention_masked(const bq_format* fp, const bq_format* fo, const void* Qq, const void* Kq, const void* Vq, void* out,
                                   int64_t B, int64_t H, int64_t S, int64_t d, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo,
                                   float score_div, int32_t causal, const uint32_t* key_mask, int64_t key_mask_words, void* stream) {
  return bq::attention_impl(fp, fo, Qq, Kq, Vq, out, B, H, S, d, ldq, ldk, ldv, ldo, score_div, (cudaStream_t)stream, causal, key_mask,
                            key_mask_words);
}

extern "C" int bq_attention_causal_q(const bq_format* fp, const bq_format* fo, const void* Qq, const void* Kq, const void* Vq,
                                     void* out_bf16, int64_t B, int64_t H, int64_t S, int64_t d, int64_t ldq, int64_t ldk,
                                     int64_t ldv, int64_t ldo, float score_div, void* stream) {
  if (!fo) return BQ_ERR_BAD_ARG;
  return bq::attention_impl(fp, fo, Qq, Kq, Vq, out_bf16, B, H, S, d, ldq, ldk, ldv, ldo, score_div, (cudaStream_t)stream);
}
