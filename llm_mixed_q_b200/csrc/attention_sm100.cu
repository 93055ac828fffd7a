// attention_sm100.cu — fused causal quantized attention for sm_100a.
//
// Replaces, for one decoder layer, the reference's chain (models/opt_quantized/modeling_opt.py:246-312,
// models/llama_quantized/modeling_llama.py:309-344):
//     scores = bmm(Qx(q), Qy(k^T))            fp32 [B*h, S, S]  written to HBM
//     scores = max(scores + causal_mask, finfo.min) ; probs = softmax(scores)       3 more HBM round trips
//     out    = bmm(Qx(probs), Qy(v))          probs quantised in 1x16 blocks along the key dim
// with ONE kernel in which scores and probabilities never leave the SM:
//     S = Q K^T on tcgen05 (bf16 operands = exact block-quantised q / k, fp32 accumulate in TMEM),
//     sweep 1: running row max / row sum of exp(s - max) (online rescaling), sweep 2: p = exp(s - max) * (1/sum)
//     (S is recomputed in the second sweep — the tensor pipe is far from being the bottleneck),
//     P quantised in registers (one thread owns a query row, so a 1x16 block is 16 consecutive registers),
//     written as a swizzled bf16 K-major smem tile and multiplied with V (MN-major operand) into a TMEM
//     accumulator.
// Blocks of P need FINAL probabilities, which is why this is a multi-sweep rather than an online-softmax
// (flash) schedule.  Key tiles above the diagonal are skipped: their probabilities are exactly 0 in the
// reference (exp(finfo.min - max) == 0) and quantise to 0.
//
// Operands are produced by bq_quantize: Qq, Kq, Vq are bf16 [B, S, h, d] (token stride given), Kq blocked
// along S (k^T's last dim), Vq along d.  d == 64.  Output: fp32 [B, S, h, d].
//
// Numerics vs the reference's torch softmax: same exp(x - max) with libdevice expf; the row sum is accumulated
// in a different order and p uses one multiplication by the correctly rounded reciprocal instead of a division
// (<= 1 ulp each).  An ulp-level difference only matters when a probability sits on a rounding boundary of the
// block format (one quantisation step there); tests/test_gpu_consumers.py states the tolerance.
//
// Warp roles (640 threads, 1 CTA/SM, persistent over (b, h, 128-row query tile) work items, heaviest first):
//   warp 0      TMA producer            warp 1   MMA issuer           warp 2   TMEM allocator
//   warps 4-19  softmax/quantise: warp w owns TMEM lane quarter (w % 4) = 32 query rows and key columns
//               [32*cq, 32*cq+32) of every 128-key tile, cq = (w - 4) / 4   (ALU-bound part: 16 warps)
#include "bq_internal.h"
#include "bq_numerics.cuh"
#include "sm100_ptx.cuh"

#include <cuda_bf16.h>
#include <math.h>

namespace bq {

constexpr int kAtBM = 128;      // query rows per work item (UMMA M)
constexpr int kAtBN = 128;      // keys per tile
constexpr int kAtD = 64;        // head dim
constexpr int kAtThreads = 640;
constexpr int kSoftmaxWarps = 16;
constexpr int kKStages = 3, kVStages = 2;
constexpr int kTileBytes = 128 * 64 * 2;           // every smem tile here is 128 rows x 128 bytes = 16 KB
constexpr int kSmemQ = 0;
constexpr int kSmemK = kSmemQ + kTileBytes;
constexpr int kSmemV = kSmemK + kKStages * kTileBytes;
constexpr int kSmemP = kSmemV + kVStages * kTileBytes;          // 2 buffers x 2 sub-tiles (64 keys each)
constexpr int kSmemX = kSmemP + 4 * kTileBytes;                 // row-stat exchange: (m, l) x 4 column quarters x 128 rows
constexpr int kSmemBar = kSmemX + 2 * 4 * 128 * 4;
constexpr int kNumBars = 2 + 2 * kKStages + 2 * kVStages + 4 + 4 + 2;
constexpr int kAtSmemBytes = kSmemBar + kNumBars * 8 + 16 + 1024;
constexpr uint32_t kAtTmemCols = 512;                            // S: 2 x 128, O: 64  -> next power of two
constexpr uint32_t kTmemO = 256;

struct AttnArgs {
  float* out;
  int B, H, S;
  int64_t ldo;          // token stride of out (elements)
  int q_tiles;          // ceil(S / 128)
  float score_div;      // scores are divided by this before the softmax (Llama: sqrt(d); OPT: 1)
  FmtParams p;          // format of P (data_in of bmm_1 / matmul_1)
};

// MN-major SWIZZLE_128B operand: rows of 128 bytes run along MN (64 bf16), 8 such rows (8 K indices) per
// 1024-byte atom.  SBO = distance between 8-K groups; LBO = distance between 64-element MN chunks (unused: N = 64).
__device__ __forceinline__ uint64_t smem_desc_sw128_mnmajor(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3fffu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
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

// Quantise 16 consecutive probabilities (one reference block) in place.
template <int KIND>
__device__ __forceinline__ void quantize_block16(float (&v)[16], const FmtParams& p) {
  uint32_t m = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) m = max(m, __float_as_uint(v[i]) & 0x7fffffffu);
  if (m == 0) return;                                  // all-zero block -> zeros (pass-through)
  const FastState fs = fast_state<KIND>(m, p);
  if (fs.ok) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = quant_elem_fast<KIND>(v[i], fs, p);
  } else {
    const BlockState st = block_state<KIND>(__uint_as_float(m), p);
#pragma unroll 1
    for (int i = 0; i < 16; ++i) v[i] = quant_elem<KIND>(v[i], st, p);
  }
}

struct Ring {
  int idx = 0;
  uint32_t phase = 0;
  __device__ __forceinline__ void advance(int n) {
    if (++idx == n) { idx = 0; phase ^= 1; }
  }
};

template <int KIND>
__global__ void __launch_bounds__(kAtThreads, 1)
attention_causal_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmV, AttnArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sb = ptx::smem_u32(smem);
  const uint32_t bar0 = sb + kSmemBar;
  // barrier map
  const uint32_t q_full = bar0, q_empty = bar0 + 8;
  auto k_full = [&](int s) { return bar0 + 8u * (2 + s); };
  auto k_empty = [&](int s) { return bar0 + 8u * (2 + kKStages + s); };
  auto v_full = [&](int s) { return bar0 + 8u * (2 + 2 * kKStages + s); };
  auto v_empty = [&](int s) { return bar0 + 8u * (2 + 2 * kKStages + kVStages + s); };
  const uint32_t bS = bar0 + 8u * (2 + 2 * kKStages + 2 * kVStages);
  auto s_full = [&](int s) { return bS + 8u * s; };
  auto s_empty = [&](int s) { return bS + 8u * (2 + s); };
  auto p_full = [&](int s) { return bS + 8u * (4 + s); };
  auto p_empty = [&](int s) { return bS + 8u * (6 + s); };
  const uint32_t o_full = bS + 8u * 8, o_empty = bS + 8u * 9;
  const uint32_t tmem_slot = bS + 8u * 10;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + kSmemBar + kNumBars * 8);
  float* xch = reinterpret_cast<float*>(smem + kSmemX);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmQ);
    ptx::prefetch_tmap(&tmK);
    ptx::prefetch_tmap(&tmV);
  }
  if (warp == 1 && lane == 0) {
    ptx::mbar_init(q_full, 1);
    ptx::mbar_init(q_empty, 1);
    for (int s = 0; s < kKStages; ++s) { ptx::mbar_init(k_full(s), 1); ptx::mbar_init(k_empty(s), 1); }
    for (int s = 0; s < kVStages; ++s) { ptx::mbar_init(v_full(s), 1); ptx::mbar_init(v_empty(s), 1); }
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
  const uint32_t tmem = *tmem_slot_ptr;

  const int items = g.B * g.H * g.q_tiles;
  // heaviest (largest query tile) first; items of one query tile are contiguous
  auto decode = [&](int w, int& b, int& h, int& qt) {
    qt = g.q_tiles - 1 - w / (g.B * g.H);
    const int r = w % (g.B * g.H);
    b = r / g.H;
    h = r % g.H;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      Ring kr, vr;
      uint32_t qphase = 0;
      for (int w = blockIdx.x; w < items; w += gridDim.x) {
        int b, h, qt;
        decode(w, b, h, qt);
        const int n = qt + 1;
        ptx::mbar_wait(q_empty, qphase ^ 1);
        ptx::mbar_expect_tx(q_full, kTileBytes);
        tma_load_4d(sb + kSmemQ, &tmQ, q_full, 0, h, qt * kAtBM, b);
        qphase ^= 1;
        for (int sweep = 0; sweep < 2; ++sweep) {
          for (int j = 0; j < n; ++j) {
            ptx::mbar_wait(k_empty(kr.idx), kr.phase ^ 1);
            ptx::mbar_expect_tx(k_full(kr.idx), kTileBytes);
            tma_load_4d(sb + kSmemK + kr.idx * kTileBytes, &tmK, k_full(kr.idx), 0, h, j * kAtBN, b);
            kr.advance(kKStages);
            if (sweep == 1) {
              ptx::mbar_wait(v_empty(vr.idx), vr.phase ^ 1);
              ptx::mbar_expect_tx(v_full(vr.idx), kTileBytes);
              tma_load_4d(sb + kSmemV + vr.idx * kTileBytes, &tmV, v_full(vr.idx), 0, h, j * kAtBN, b);
              vr.advance(kVStages);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idS = ptx::idesc_bf16_f32(kAtBM, kAtBN);
      constexpr uint32_t idO = idesc_bf16_f32_bmn(kAtBM, kAtD);
      Ring kr, vr, sr, pr;
      uint32_t qphase = 0, ophase = 0;
      const uint64_t qdesc = ptx::smem_desc_sw128_kmajor(sb + kSmemQ);
      auto issue_S = [&]() {
        ptx::mbar_wait(k_full(kr.idx), kr.phase);
        ptx::mbar_wait(s_empty(sr.idx), sr.phase ^ 1);
        ptx::tc_fence_after();
        const uint64_t kdesc = ptx::smem_desc_sw128_kmajor(sb + kSmemK + kr.idx * kTileBytes);
#pragma unroll
        for (int k = 0; k < kAtD / 16; ++k)
          ptx::umma_bf16(tmem + (uint32_t)(sr.idx * kAtBN), qdesc + (uint64_t)(2 * k), kdesc + (uint64_t)(2 * k), idS, k != 0);
        ptx::umma_commit(k_empty(kr.idx));
        ptx::umma_commit(s_full(sr.idx));
        kr.advance(kKStages);
        sr.advance(2);
      };
      for (int w = blockIdx.x; w < items; w += gridDim.x) {
        int b, h, qt;
        decode(w, b, h, qt);
        const int n = qt + 1;
        ptx::mbar_wait(q_full, qphase);
        qphase ^= 1;
        for (int j = 0; j < n; ++j) issue_S();                // statistics sweep
        issue_S();                                            // S(0) of the final sweep
        for (int j = 0; j < n; ++j) {
          if (j + 1 < n) issue_S();                           // keep the softmax warps one tile ahead
          ptx::mbar_wait(v_full(vr.idx), vr.phase);
          ptx::mbar_wait(p_full(pr.idx), pr.phase);
          if (j == 0) { ptx::mbar_wait(o_empty, ophase ^ 1); }
          ptx::tc_fence_after();
          const uint32_t pbase = sb + kSmemP + pr.idx * 2 * kTileBytes;
          const uint32_t vbase = sb + kSmemV + vr.idx * kTileBytes;
#pragma unroll
          for (int k = 0; k < kAtBN / 16; ++k) {
            const uint64_t adesc = ptx::smem_desc_sw128_kmajor(pbase + (k >> 2) * kTileBytes) + (uint64_t)(2 * (k & 3));
            const uint64_t bdesc = smem_desc_sw128_mnmajor(vbase + k * 16 * 128);
            ptx::umma_bf16(tmem + kTmemO, adesc, bdesc, idO, (j | k) != 0);
          }
          ptx::umma_commit(v_empty(vr.idx));
          ptx::umma_commit(p_empty(pr.idx));
          vr.advance(kVStages);
          pr.advance(2);
        }
        ptx::umma_commit(o_full);
        ptx::umma_commit(q_empty);
        ophase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax / quantise / epilogue
    const int quarter = warp & 3;                 // TMEM lane quarter
    const int cq = (warp - 4) >> 2;               // which 32 key columns of every tile
    const int r_in = quarter * 32 + lane;         // query row inside the tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    Ring sr, pr;
    uint32_t ophase = 0;
    const bool scale = (g.score_div != 1.0f);
    float* xm = xch;                              // [4][128] partial maxima
    float* xl = xch + 4 * 128;                    // [4][128] partial sums
    for (int w = blockIdx.x; w < items; w += gridDim.x) {
      int b, h, qt;
      decode(w, b, h, qt);
      const int n = qt + 1;
      const int row = qt * kAtBM + r_in;
      float m = -INFINITY, l = 0.f, inv_l = 0.f;
      for (int sweep = 0; sweep < 2; ++sweep) {
        for (int j = 0; j < n; ++j) {
          ptx::mbar_wait(s_full(sr.idx), sr.phase);
          ptx::tc_fence_after();
          uint32_t r[32];
          ptx::tmem_ld_32x32(tmem + lane_addr + (uint32_t)(sr.idx * kAtBN + cq * 32), r);
          ptx::tmem_ld_wait();
          const int cb = j * kAtBN + cq * 32;                 // first key column of this thread's slice
          const int nvalid = (j == n - 1) ? min(max(row - cb + 1, 0), 32) : 32;   // causal: keys <= row
          if (scale) {
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__fdiv_rn(__uint_as_float(r[i]), g.score_div));
          }
          if (sweep == 0) {
            float tmax = -INFINITY;
#pragma unroll
            for (int i = 0; i < 32; ++i) tmax = (i < nvalid) ? fmaxf(tmax, __uint_as_float(r[i])) : tmax;
            if (tmax > m) {                                   // online rescale of the running sum
              l = __fmul_rn(l, expf(__fsub_rn(m, tmax)));
              m = tmax;
            }
            float acc = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) acc = (i < nvalid) ? __fadd_rn(acc, expf(__fsub_rn(__uint_as_float(r[i]), m))) : acc;
            l = __fadd_rn(l, acc);
          } else {
            ptx::mbar_wait(p_empty(pr.idx), pr.phase ^ 1);
            // probabilities of 32 keys = two reference blocks; quantise and store as bf16 into the P tile
            uint8_t* prow = smem + kSmemP + (pr.idx * 2 + (cq >> 1)) * kTileBytes + r_in * 128;
#pragma unroll
            for (int blk = 0; blk < 2; ++blk) {
              float v[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float e = expf(__fsub_rn(__uint_as_float(r[blk * 16 + i]), m));
                v[i] = (blk * 16 + i < nvalid) ? __fmul_rn(e, inv_l) : 0.f;
              }
              quantize_block16<KIND>(v, g.p);
              uint32_t w32[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                __nv_bfloat162 t2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
                w32[i] = *reinterpret_cast<uint32_t*>(&t2);
              }
              // 16-byte chunk index inside the 128-byte row, XOR-swizzled by (row & 7)  (SWIZZLE_128B)
              const int chunk = (cq & 1) * 4 + blk * 2;
              *reinterpret_cast<uint4*>(prow + ((chunk ^ (r_in & 7)) << 4)) = make_uint4(w32[0], w32[1], w32[2], w32[3]);
              *reinterpret_cast<uint4*>(prow + (((chunk + 1) ^ (r_in & 7)) << 4)) = make_uint4(w32[4], w32[5], w32[6], w32[7]);
            }
            ptx::fence_proxy_async_smem();     // generic-proxy smem writes -> visible to the MMA (async proxy)
          }
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (sweep == 1) ptx::mbar_arrive(p_full(pr.idx));
            ptx::mbar_arrive(s_empty(sr.idx));
          }
          if (sweep == 1) pr.advance(2);
          sr.advance(2);
        }
        if (sweep == 0) {
          // merge the four column quarters of every row:  m = max m_c,  l = sum_c l_c * exp(m_c - m)
          xm[cq * 128 + r_in] = m;
          xl[cq * 128 + r_in] = l;
          named_bar_sync(1, kSoftmaxWarps * 32);
          float mm = xm[r_in];
#pragma unroll
          for (int c = 1; c < 4; ++c) mm = fmaxf(mm, xm[c * 128 + r_in]);
          float ll = 0.f;
#pragma unroll
          for (int c = 0; c < 4; ++c) ll = __fadd_rn(ll, __fmul_rn(xl[c * 128 + r_in], expf(__fsub_rn(xm[c * 128 + r_in], mm))));
          m = mm;
          l = ll;
          inv_l = __frcp_rn(l);
          named_bar_sync(1, kSoftmaxWarps * 32);
        }
      }
      // ---- epilogue: O (128 x 64 fp32 in TMEM) -> global; this warp owns 16 of the 64 columns of its 32 rows
      ptx::mbar_wait(o_full, ophase);
      ophase ^= 1;
      ptx::tc_fence_after();
      uint32_t r[16];
      ptx::tmem_ld_32x16(tmem + lane_addr + kTmemO + (uint32_t)(cq * 16), r);
      ptx::tmem_ld_wait();
      if (row < g.S) {
        float* o = g.out + ((int64_t)b * g.S + row) * g.ldo + (int64_t)h * kAtD + cq * 16;
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          *reinterpret_cast<float4*>(o + i) = make_float4(__uint_as_float(r[i]), __uint_as_float(r[i + 1]),
                                                          __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(o_empty);
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
// host side
// ------------------------------------------------------------------------------------------------
int make_params(const bq_format* f, FmtParams* p);
int make_tmap_bf16_4d(CUtensorMap* tm, const void* base, int64_t d, int64_t S, int64_t H, int64_t B, int64_t ld_tok,
                      int box_rows);

template <int KIND>
static int launch_attention(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnArgs& g,
                            cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    BQ_CUDA_CHECK(cudaFuncSetAttribute(attention_causal_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtSmemBytes));
    attr = true;
  }
  const int items = g.B * g.H * g.q_tiles;
  const int grid = std::min(items, num_sms());
  {
    LaunchScope ls(kKernAttention, st);
    attention_causal_kernel<KIND><<<grid, kAtThreads, kAtSmemBytes, st>>>(tq, tk, tv, g);
  }
  BQ_CUDA_CHECK(cudaGetLastError());
  return BQ_OK;
}

}  // namespace bq

extern "C" int bq_attention_causal(const bq_format* fp, const void* Qq, const void* Kq, const void* Vq, float* out,
                                   int64_t B, int64_t H, int64_t S, int64_t d, int64_t ldq, int64_t ldk, int64_t ldv,
                                   int64_t ldo, float score_div, void* stream) {
  using namespace bq;
  if (!fp || B < 0 || H < 0 || S < 0) return BQ_ERR_BAD_ARG;
  if (B == 0 || H == 0 || S == 0) return BQ_OK;
  if (!Qq || !Kq || !Vq || !out) return BQ_ERR_BAD_ARG;
  if (d != kAtD) return BQ_ERR_UNSUPPORTED;
  if (fp->kind != BQ_KIND_BLOCK_FP && fp->kind != BQ_KIND_BLOCK_MINIFLOAT) return BQ_ERR_UNSUPPORTED;
  if (fp->block_rows != 1 || fp->block_cols != 16) return BQ_ERR_UNSUPPORTED;
  if ((ldq % 8) || (ldk % 8) || (ldv % 8) || (ldo % 4) || ((uintptr_t)Qq % 16) || ((uintptr_t)Kq % 16) ||
      ((uintptr_t)Vq % 16) || ((uintptr_t)out % 16))
    return BQ_ERR_BAD_ARG;
  if (B * H * ((S + 127) / 128) > 0x7fffffffll || S > 0x7fffffff) return BQ_ERR_UNSUPPORTED;
  AttnArgs g;
  int rc = make_params(fp, &g.p);
  if (rc) return rc;
  g.p.fold_zero = 0;
  g.out = out; g.B = (int)B; g.H = (int)H; g.S = (int)S; g.ldo = ldo;
  g.q_tiles = (int)((S + kAtBM - 1) / kAtBM);
  g.score_div = score_div;
  CUtensorMap tq, tk, tv;
  if ((rc = make_tmap_bf16_4d(&tq, Qq, d, S, H, B, ldq, kAtBM))) return rc;
  if ((rc = make_tmap_bf16_4d(&tk, Kq, d, S, H, B, ldk, kAtBN))) return rc;
  if ((rc = make_tmap_bf16_4d(&tv, Vq, d, S, H, B, ldv, kAtBN))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (fp->kind == BQ_KIND_BLOCK_FP) return launch_attention<kBlockFP>(tq, tk, tv, g, st);
  return launch_attention<kBlockMinifloat>(tq, tk, tv, g, st);
}
