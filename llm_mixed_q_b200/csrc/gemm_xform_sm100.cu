// gemm_xform_sm100.cu — tcgen05 GEMM whose mainloop TRANSFORMS one operand in shared memory before the MMA reads it.
//
// Two north-star variants of the quantized Linear (reference quantized_modules/linear.py:59-76: y = F.linear(Qx(x), Qw(W), Qb(b)))
// that the two-launch path (quantize kernel -> bf16 GEMM, ops.cu: bq_linear) does not cover:
//
//   XF_QUANT_A   activation quantisation in the GEMM PROLOGUE: A arrives as raw fp32 [M][K] tiles (TMA, 128B swizzle), eight
//                transform warps apply the block quantizer (1x16 blocks along K, bq_numerics.cuh — same arithmetic as bq_quantize)
//                and write the exact quantised values as a bf16 K-major SWIZZLE_128B tile that tcgen05.mma consumes.  One launch,
//                no bf16 copy of x in HBM; the price is that every N-tile of a row block re-quantises the same A tile.
//   XF_PACKED_B  weights stay PACKED in HBM — per 256 K-elements of a row: 32*w bytes of sign+magnitude fields (w bits each, the
//                reference's `width`) followed by 16 shared-exponent bytes (one per block of 16) = w + 0.5 bits per element, the
//                reference's own cost model (quantized_layer_profiler.py:18-27) — and are decoded to bf16 in the mainloop.
//                bf16 carries 16 bits per element: at small M (weights dominate HBM traffic) the packed stream is 2.46x (W6) to
//                3.6x (W4) fewer bytes.
//
// Warp roles (512 threads, 1 CTA / SM, persistent):
//   warp 0  TMA producer of the DIRECT operand (bf16 tiles straight into the MMA ring)      warp 1  MMA issuer
//   warp 2  TMEM allocator                     warp 3  TMA producer of the RAW operand (fp32 A tiles / packed B groups)
//   warps 4-7  epilogue (TMEM -> registers -> +bias -> fp32 global)          warps 8-15  transform (raw ring -> MMA ring)
// Rings: raw (raw_full: TMA bytes; raw_empty: 8 transform warps) and MMA (op_full: 1 producer arrive + TMA bytes + 8 transform
// warps; op_empty: tcgen05.commit), TMEM full / empty x 2 accumulators.
#include "bq_internal.h"
#include "bq_blockops.cuh"
#include "sm100_ptx.cuh"

namespace bq {

constexpr int kXBM = 128, kXBK = 64;
constexpr int kXformWarps = 8;
constexpr int kXThreads = 512;
enum { XF_QUANT_A = 1, XF_PACKED_B = 2, XF_PACKED_A = 3 };

template <int BN, int XF> struct XCfg {
  static constexpr int kStageA = kXBM * kXBK * 2;                  // 16 KB
  static constexpr int kStageB = BN * kXBK * 2;
  static constexpr int kStage = kStageA + kStageB;
  static constexpr int kStages = XF == XF_QUANT_A ? 3 : (XF == XF_PACKED_B && BN <= 32 ? 8 : 4);
  // raw stage: XF_QUANT_A: one 64-wide K tile of fp32 A = two {32 floats x 128 rows} swizzled boxes; XF_PACKED_B: one 256-element K
  // group of BN packed rows, sized for the widest format (w = 8: 272 bytes per row)
  static constexpr int kRawRows = XF == XF_PACKED_A ? kXBM : BN;      // packed rows per raw stage
  static constexpr int kRawStage = XF == XF_QUANT_A ? kXBM * kXBK * 4 : ((kRawRows * 272 + 1023) / 1024) * 1024;
  static constexpr int kRawStages = XF == XF_QUANT_A ? 2 : (XF == XF_PACKED_B && BN <= 32 ? 6 : (XF == XF_PACKED_A && BN <= 32 ? 4 : 2));   // a packed raw stage feeds four MMA stages
  static constexpr int kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;
  static constexpr int kOffRaw = kStages * kStage;
  static constexpr int kOffBar = kOffRaw + kRawStages * kRawStage;
  static constexpr int kNumBars = 2 * kStages + 2 * kRawStages + 4;
  static constexpr int kSmemBytes = kOffBar + kNumBars * 8 + 16 + 1024;
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
};

struct XArgs {
  float* C;
  const float* bias;
  int M, N, K;
  int64_t ldc;
  int tiles_m, tiles_n;
  FmtParams q;                 // XF_QUANT_A: the x-quantizer
  int w;                       // XF_PACKED_B: field width (2..8)
  int exp_bias;                //              exponent byte = E + exp_bias
  int row_bytes;               //              32*w + 16
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_v4u(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t bf16x2_fma_u(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// 32 packed fields of W bits (sign in the top bit of a field, magnitude below) + the two blocks' scales -> 16 bf16x2 words.
// value = (-1)^sign * magnitude * 2^(E - m): 0x4300 | magnitude is the bf16 encoding of 128 + magnitude (magnitude < 128), and
// fma(128 + magnitude, s, -128 s) is exact, so the decode is one packed FMA per element pair plus the field extraction.
template <int W>
__device__ __forceinline__ void decode_unit(const uint32_t (&words)[8], uint32_t s0, uint32_t s1, uint32_t (&out)[16]) {
  constexpr uint32_t kMag = (1u << (W - 1)) - 1u;
  // -128 * s as bf16x2: flip the sign and add 7 to the exponent field (s is a normal power of two or zero)
  const uint32_t nb0 = s0 ? ((s0 + 0x03800380u) | 0x80008000u) : 0u, nb1 = s1 ? ((s1 + 0x03800380u) | 0x80008000u) : 0u;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int e0 = 2 * i, e1 = 2 * i + 1;
    const int b0 = e0 * W, b1 = e1 * W;
    uint32_t f0 = words[b0 >> 5] >> (b0 & 31);
    if ((b0 & 31) + W > 32) f0 |= words[(b0 >> 5) + 1] << (32 - (b0 & 31));
    uint32_t f1 = words[b1 >> 5] >> (b1 & 31);
    if ((b1 & 31) + W > 32) f1 |= words[(b1 >> 5) + 1] << (32 - (b1 & 31));
    const uint32_t mag = (f0 & kMag) | ((f1 & kMag) << 16);
    const uint32_t sgn = (((f0 >> (W - 1)) & 1u) << 15) | (((f1 >> (W - 1)) & 1u) << 31);
    const uint32_t s = i < 8 ? s0 : s1, nb = i < 8 ? nb0 : nb1;
    out[i] = bf16x2_fma_u(mag | 0x43004300u, s, nb) | sgn;
  }
}
// scale 2^(E - m) of a block as a bf16x2 pair (both halves equal); exponents below the bf16 normal range flush to zero
__device__ __forceinline__ uint32_t scale_bf16x2(int ebyte, int exp_bias, int m) {
  const int f = ebyte - exp_bias - m + 127;
  const uint32_t h = (f >= 1 && f <= 254) ? ((uint32_t)f << 7) : 0u;
  return h | (h << 16);
}

template <int BN, int XF>
__global__ void __launch_bounds__(kXThreads, 1)
gemm_xform_kernel(const __grid_constant__ CUtensorMap tmDirect, const __grid_constant__ CUtensorMap tmRaw, XArgs g) {
  using Cfg = XCfg<BN, XF>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + Cfg::kOffBar;
  auto op_full = [&](int s) { return bar_base + 8u * s; };
  auto op_empty = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
  const uint32_t bR = bar_base + 8u * (2 * Cfg::kStages);
  auto raw_full = [&](int s) { return bR + 8u * s; };
  auto raw_empty = [&](int s) { return bR + 8u * (Cfg::kRawStages + s); };
  const uint32_t bT = bR + 8u * (2 * Cfg::kRawStages);
  auto tfull = [&](int a) { return bT + 8u * a; };
  auto tempty = [&](int a) { return bT + 8u * (2 + a); };
  const uint32_t tmem_slot = bT + 8u * 4;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmDirect);
    ptx::prefetch_tmap(&tmRaw);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      ptx::mbar_init(op_full(s), 1 + kXformWarps);
      ptx::mbar_init(op_empty(s), 1);
    }
    for (int s = 0; s < Cfg::kRawStages; ++s) {
      ptx::mbar_init(raw_full(s), 1);
      ptx::mbar_init(raw_empty(s), kXformWarps);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(tfull(a), 1);
      ptx::mbar_init(tempty(a), 4);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  const int num_kb = g.K / kXBK;                       // K % 64 == 0 (XF_PACKED_B: K % 256 == 0) — checked on the host
  const int total_tiles = g.tiles_m * g.tiles_n;
  auto decode = [&](int tile, int& row0, int& nb) {
    const int mb = tile / g.tiles_n;                   // n fastest: the CTAs of a wave share A rows through L2
    nb = tile - mb * g.tiles_n;                        // XF_PACKED_A: tiles_m == 1, nb counts 128-row blocks of the weight
    row0 = mb * kXBM;
  };

  if (warp == 0) {
    // ---------------------------------------------------------------- direct operand: bf16 tiles straight into the MMA ring
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int row0, nb;
        decode(tile, row0, nb);
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(op_empty(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * Cfg::kStage;
          if (XF == XF_QUANT_A) {
            ptx::mbar_expect_tx(op_full(stage), Cfg::kStageB);
            ptx::tma_load_3d(sa + Cfg::kStageA, &tmDirect, op_full(stage), kb * kXBK, nb * BN, 0);
          } else if (XF == XF_PACKED_A) {                        // swapped roles: the activation (<= BN rows) is the B operand
            ptx::mbar_expect_tx(op_full(stage), Cfg::kStageB);
            ptx::tma_load_3d(sa + Cfg::kStageA, &tmDirect, op_full(stage), kb * kXBK, 0, 0);
          } else {
            ptx::mbar_expect_tx(op_full(stage), Cfg::kStageA);
            ptx::tma_load_3d(sa, &tmDirect, op_full(stage), kb * kXBK, row0, 0);
          }
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 3) {
    // ---------------------------------------------------------------- raw operand: fp32 A tiles / packed B groups
    if (lane == 0) {
      int rs = 0;
      uint32_t rphase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int row0, nb;
        decode(tile, row0, nb);
        const int n_raw = XF == XF_QUANT_A ? num_kb : num_kb / 4;
        for (int i = 0; i < n_raw; ++i) {
          ptx::mbar_wait(raw_empty(rs), rphase ^ 1);
          const uint32_t dst = smem_base + Cfg::kOffRaw + rs * Cfg::kRawStage;
          if (XF == XF_QUANT_A) {
            ptx::mbar_expect_tx(raw_full(rs), kXBM * kXBK * 4);
            tma_load_2d(dst, &tmRaw, raw_full(rs), i * kXBK, row0);                    // k [0,32) of the tile: 128 rows x 128 bytes
            tma_load_2d(dst + kXBM * 128, &tmRaw, raw_full(rs), i * kXBK + 32, row0);  // k [32,64)
          } else {
            ptx::mbar_expect_tx(raw_full(rs), (uint32_t)(Cfg::kRawRows * g.row_bytes));
            tma_load_2d(dst, &tmRaw, raw_full(rs), i * (g.row_bytes / 4), nb * Cfg::kRawRows);    // one 256-element group of packed rows
          }
          if (++rs == Cfg::kRawStages) { rs = 0; rphase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      const uint32_t idesc = ptx::idesc_bf16_f32(kXBM, BN);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        ptx::mbar_wait(tempty(acc), acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(op_full(stage), phase);
          ptx::tc_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::kStage;
          const uint64_t adesc = ptx::smem_desc_sw128_kmajor(sa);
          const uint64_t bdesc = ptx::smem_desc_sw128_kmajor(sa + Cfg::kStageA);
#pragma unroll
          for (int k = 0; k < kXBK / 16; ++k)
            ptx::umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
          ptx::umma_commit(op_empty(stage));
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit(tfull(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 8) {
    // ---------------------------------------------------------------- transform warps: raw ring -> MMA ring
    const int t = threadIdx.x - 256;                 // 0..255
    int stage = 0, rs = 0;
    uint32_t phase = 0, rphase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      if (XF == XF_QUANT_A) {
        // thread -> (row, 32-float half of the 64-wide K tile) = two quantiser blocks = one 128-byte swizzled row of a raw box
        const int row = t >> 1, half = t & 1, sw = row & 7;
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(raw_full(rs), rphase);
          const uint32_t src = smem_base + Cfg::kOffRaw + rs * Cfg::kRawStage + half * (kXBM * 128) + row * 128;
          float v0[16], v1[16];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 q4 = lds_v4(src + ((c ^ sw) << 4));
            float* d = c < 4 ? &v0[4 * c] : &v1[4 * (c - 4)];
            d[0] = u2f(q4.x); d[1] = u2f(q4.y); d[2] = u2f(q4.z); d[3] = u2f(q4.w);
          }
          // Cross-proxy write-after-read: the slot is about to be refilled by TMA (async proxy) while it was read through the generic
          // proxy.  Without a proxy fence between the reads and the release, ~1e-4 of the row blocks saw the NEXT K tile's data
          // (measured: non-deterministic outputs, gone with this fence and only with it; a fence after the full-barrier wait is not
          // needed).  The mbarrier arrive alone orders the reads within the generic proxy only.
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(raw_empty(rs));            // the raw tile is in registers
          if (++rs == Cfg::kRawStages) { rs = 0; rphase ^= 1; }
          quantize_signed16_rt(v0, g.q);
          quantize_signed16_rt(v1, g.q);
          ptx::mbar_wait(op_empty(stage), phase ^ 1);                // the MMAs that read this slot have retired
          const uint32_t dst = smem_base + stage * Cfg::kStage + row * 128;
          const int c0 = half * 4;
          sts_v4u(dst + (((c0 + 0) ^ sw) << 4), pack_bf16_rn(v0[0], v0[1]), pack_bf16_rn(v0[2], v0[3]), pack_bf16_rn(v0[4], v0[5]), pack_bf16_rn(v0[6], v0[7]));
          sts_v4u(dst + (((c0 + 1) ^ sw) << 4), pack_bf16_rn(v0[8], v0[9]), pack_bf16_rn(v0[10], v0[11]), pack_bf16_rn(v0[12], v0[13]), pack_bf16_rn(v0[14], v0[15]));
          sts_v4u(dst + (((c0 + 2) ^ sw) << 4), pack_bf16_rn(v1[0], v1[1]), pack_bf16_rn(v1[2], v1[3]), pack_bf16_rn(v1[4], v1[5]), pack_bf16_rn(v1[6], v1[7]));
          sts_v4u(dst + (((c0 + 3) ^ sw) << 4), pack_bf16_rn(v1[8], v1[9]), pack_bf16_rn(v1[10], v1[11]), pack_bf16_rn(v1[12], v1[13]), pack_bf16_rn(v1[14], v1[15]));
          ptx::fence_proxy_async_smem();                             // generic-proxy writes -> visible to the MMA (async proxy)
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(op_full(stage));
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      } else if (XF == XF_PACKED_B && BN == 32) {
        // decode regime (M <= 128, the weight stream is the bound): the 256 threads take ONE 256-element group of the 32 packed rows
        // at a time — thread -> (row, unit of 32 elements) — and fill FOUR MMA stages per step, so the per-stage hand-over latency
        // (wait empty -> store -> proxy fence -> arrive) is paid once per 256 K-elements instead of once per 64
        static_assert(BN != 32 || XF != XF_PACKED_B || Cfg::kStages % 4 == 0, "four MMA stages are filled per packed group");
        const int row = t >> 3, unit = t & 7, sub = unit >> 1, u01 = unit & 1, sw = row & 7;
        const int m = g.w - 1;
        for (int grp = 0; grp < num_kb / 4; ++grp) {
          ptx::mbar_wait(raw_full(rs), rphase);
          const uint32_t rbase = smem_base + Cfg::kOffRaw + rs * Cfg::kRawStage + row * g.row_bytes;
          uint32_t words[8], outw[16];
#pragma unroll
          for (int i = 0; i < 8; ++i) words[i] = i < g.w ? lds_u32(rbase + unit * 4 * g.w + 4 * i) : 0u;
          const uint32_t eb = lds_u32(rbase + 32 * g.w + (unit >> 1) * 4);
          const int sh = (unit & 1) * 16;
          const uint32_t s0 = scale_bf16x2((int)((eb >> sh) & 0xffu), g.exp_bias, m);
          const uint32_t s1 = scale_bf16x2((int)((eb >> (sh + 8)) & 0xffu), g.exp_bias, m);
          switch (g.w) {
            case 2: decode_unit<2>(words, s0, s1, outw); break;
            case 3: decode_unit<3>(words, s0, s1, outw); break;
            case 4: decode_unit<4>(words, s0, s1, outw); break;
            case 5: decode_unit<5>(words, s0, s1, outw); break;
            case 6: decode_unit<6>(words, s0, s1, outw); break;
            case 7: decode_unit<7>(words, s0, s1, outw); break;
            default: decode_unit<8>(words, s0, s1, outw); break;
          }
          // stages stage .. stage + 3 (kStages % 4 == 0: no wrap inside a step); every warp holds all eight units of four rows, so it
          // contributes to — and arrives on — each of the four stages
#pragma unroll
          for (int j = 0; j < 4; ++j) ptx::mbar_wait(op_empty(stage + j), phase ^ 1);
          const uint32_t dst = smem_base + (stage + sub) * Cfg::kStage + Cfg::kStageA + row * 128;
          const int c0 = u01 * 4;
#pragma unroll
          for (int c = 0; c < 4; ++c) sts_v4u(dst + (((c0 + c) ^ sw) << 4), outw[4 * c], outw[4 * c + 1], outw[4 * c + 2], outw[4 * c + 3]);
          ptx::fence_proxy_async_smem();                 // also orders this thread's raw-slot reads before the slot's release
          __syncwarp();
          if (lane == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) ptx::mbar_arrive(op_full(stage + j));
            ptx::mbar_arrive(raw_empty(rs));
          }
          stage += 4;
          if (stage == Cfg::kStages) { stage = 0; phase ^= 1; }
          if (++rs == Cfg::kRawStages) { rs = 0; rphase ^= 1; }
        }
      } else if (XF == XF_PACKED_A) {
        // decode regime with swapped roles: thread -> (weight row, left / right 32-element unit) of each of the four 64-wide K tiles
        // of a 256-element group.  All four units are decoded into registers FIRST (ALU work that overlaps the MMAs still reading the
        // four stages), then the stages are claimed, filled and handed over with ONE proxy fence per group — the per-stage hand-over
        // (wait -> store -> fence -> arrive, ~0.7 us) had made the kernel latency-bound at 1.2 TB/s of packed bytes.
        static_assert(XF != XF_PACKED_A || Cfg::kStages == 4, "one packed group fills the whole MMA ring");
        const int row = t >> 1, u01 = t & 1, sw = row & 7;
        const int m = g.w - 1;
        for (int grp = 0; grp < num_kb / 4; ++grp) {
          ptx::mbar_wait(raw_full(rs), rphase);
          const uint32_t rbase = smem_base + Cfg::kOffRaw + rs * Cfg::kRawStage + row * g.row_bytes;
          // the four stages of the ring are claimed up front (the MMAs of the previous group are short: 16 instructions), then one
          // SMALL loop body decodes and stores a unit per trip — unrolling it four times over seven field widths put 200 KB of code
          // in front of a 32 KB instruction cache and the kernel ran at a third of this speed
#pragma unroll
          for (int sub = 0; sub < 4; ++sub) ptx::mbar_wait(op_empty(sub), phase ^ 1);
          const int c0 = u01 * 4;
#pragma unroll 1
          for (int sub = 0; sub < 4; ++sub) {
            const int unit = sub * 2 + u01;
            uint32_t words[8], outw[16];
#pragma unroll
            for (int i = 0; i < 8; ++i) words[i] = i < g.w ? lds_u32(rbase + unit * 4 * g.w + 4 * i) : 0u;
            const uint32_t eb = lds_u32(rbase + 32 * g.w + (unit >> 1) * 4);
            const int sh = (unit & 1) * 16;
            const uint32_t s0 = scale_bf16x2((int)((eb >> sh) & 0xffu), g.exp_bias, m);
            const uint32_t s1 = scale_bf16x2((int)((eb >> (sh + 8)) & 0xffu), g.exp_bias, m);
            switch (g.w) {
              case 2: decode_unit<2>(words, s0, s1, outw); break;
              case 3: decode_unit<3>(words, s0, s1, outw); break;
              case 4: decode_unit<4>(words, s0, s1, outw); break;
              case 5: decode_unit<5>(words, s0, s1, outw); break;
              case 6: decode_unit<6>(words, s0, s1, outw); break;
              case 7: decode_unit<7>(words, s0, s1, outw); break;
              default: decode_unit<8>(words, s0, s1, outw); break;
            }
            const uint32_t dst = smem_base + sub * Cfg::kStage + row * 128;
#pragma unroll
            for (int c = 0; c < 4; ++c) sts_v4u(dst + (((c0 + c) ^ sw) << 4), outw[4 * c], outw[4 * c + 1], outw[4 * c + 2], outw[4 * c + 3]);
          }
          ptx::fence_proxy_async_smem();                 // also orders this thread's raw-slot reads before the slot's release
          __syncwarp();
          if (lane == 0) {
#pragma unroll
            for (int sub = 0; sub < 4; ++sub) ptx::mbar_arrive(op_full(sub));
            ptx::mbar_arrive(raw_empty(rs));
          }
          phase ^= 1;
          if (++rs == Cfg::kRawStages) { rs = 0; rphase ^= 1; }
        }
      } else {
        // thread -> (row, 32-element unit of the 64-wide K tile)
        const int row = t >> 1, u01 = t & 1, sw = row & 7;
        const bool active = row < Cfg::kRawRows;
        constexpr int kDstOff = Cfg::kStageA;
        const int m = g.w - 1;
        for (int grp = 0; grp < num_kb / 4; ++grp) {
          ptx::mbar_wait(raw_full(rs), rphase);
          const uint32_t rbase = smem_base + Cfg::kOffRaw + rs * Cfg::kRawStage + (active ? row : 0) * g.row_bytes;
#pragma unroll 1
          for (int sub = 0; sub < 4; ++sub) {
            uint32_t outw[16];
            if (active) {
              const int unit = sub * 2 + u01;                        // 8 units of 32 elements per 256-element group
              uint32_t words[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) words[i] = i < g.w ? lds_u32(rbase + unit * 4 * g.w + 4 * i) : 0u;
              const uint32_t eb = lds_u32(rbase + 32 * g.w + (unit >> 1) * 4);      // 4 exponent bytes: blocks 4*(unit/2) .. +3
              const int sh = (unit & 1) * 16;
              const uint32_t s0 = scale_bf16x2((int)((eb >> sh) & 0xffu), g.exp_bias, m);
              const uint32_t s1 = scale_bf16x2((int)((eb >> (sh + 8)) & 0xffu), g.exp_bias, m);
              switch (g.w) {
                case 2: decode_unit<2>(words, s0, s1, outw); break;
                case 3: decode_unit<3>(words, s0, s1, outw); break;
                case 4: decode_unit<4>(words, s0, s1, outw); break;
                case 5: decode_unit<5>(words, s0, s1, outw); break;
                case 6: decode_unit<6>(words, s0, s1, outw); break;
                case 7: decode_unit<7>(words, s0, s1, outw); break;
                default: decode_unit<8>(words, s0, s1, outw); break;
              }
            }
            ptx::mbar_wait(op_empty(stage), phase ^ 1);
            if (active) {
              const uint32_t dst = smem_base + stage * Cfg::kStage + kDstOff + row * 128;
              const int c0 = u01 * 4;
#pragma unroll
              for (int c = 0; c < 4; ++c)
                sts_v4u(dst + (((c0 + c) ^ sw) << 4), outw[4 * c], outw[4 * c + 1], outw[4 * c + 2], outw[4 * c + 3]);
            }
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(op_full(stage));
            if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
          }
          // (the generic-proxy reads of this group are already behind the proxy fence of the last sub-tile's op_full hand-over)
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(raw_empty(rs));
          if (++rs == Cfg::kRawStages) { rs = 0; rphase ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- epilogue: TMEM -> registers -> (+bias) -> fp32 global
    const int q = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int row0, nb;
      decode(tile, row0, nb);
      ptx::mbar_wait(tfull(acc), acc_phase);
      ptx::tc_fence_after();
      const int row = row0 + q * 32 + lane;
      float* crow = g.C + (int64_t)row * g.ldc;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c * 32), r);
        ptx::tmem_ld_wait();
        if (XF == XF_PACKED_A) {
          // accumulator lane = weight row n, column = activation row m: y[m][n] — for a fixed m the 32 lanes of the warp write 32
          // consecutive n (one 128-byte segment)
          const int n = nb * kXBM + q * 32 + lane;
          const float bv = (g.bias && n < g.N) ? g.bias[n] : 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int mrow = c * 32 + j;
            if (mrow < g.M && n < g.N) g.C[(int64_t)mrow * g.ldc + n] = __fadd_rn(u2f(r[j]), bv);
          }
          continue;
        }
        const int col0 = nb * BN + c * 32;
        if (row < g.M && col0 < g.N) {                             // N % 32 == 0, ldc % 4 == 0, pointers 16-byte aligned: host-checked
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 o = make_float4(u2f(r[j]), u2f(r[j + 1]), u2f(r[j + 2]), u2f(r[j + 3]));
            if (g.bias) {
              const float4 bv = *reinterpret_cast<const float4*>(g.bias + col0 + j);
              o.x = __fadd_rn(o.x, bv.x); o.y = __fadd_rn(o.y, bv.y); o.z = __fadd_rn(o.z, bv.z); o.w = __fadd_rn(o.w, bv.w);
            }
            *reinterpret_cast<float4*>(crow + col0 + j) = o;
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(tempty(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// packing: quantised fp32 weights -> sign+magnitude fields + one exponent byte per block of 16
// ------------------------------------------------------------------------------------------------
struct PackArgs {
  const float* W;
  uint8_t* packed;
  unsigned long long* mismatches;
  int64_t N, K, ldw, row_bytes;
  int w, exp_bias, e_lo, e_hi;
};
// one thread per block of 16 values (already on the block_fp grid: the PTQ overwrite of linear.py:66-70 ran).  E = exponent of the
// block's largest magnitude (clamped to the format's range); every value is an integer multiple of 2^(E-m) below 2^m.  A value that
// is not (a pass-through element |x| <= 1e-8, which the reference leaves unquantised) is rounded to the grid and COUNTED.
__global__ void __launch_bounds__(256) pack_weight_kernel(PackArgs a) {
  const int64_t bpr = a.K / 16, nblk = a.N * bpr;
  const int m = a.w - 1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nblk; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / bpr, b = i - n * bpr;
    const float* src = a.W + n * a.ldw + b * 16;
    float v[16];
    uint32_t mx = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 q = *reinterpret_cast<const float4*>(src + 4 * j);
      v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) mx = max(mx, __float_as_uint(v[j]) & 0x7fffffffu);
    int E = a.e_lo;
    if (mx >= 0x00800000u) E = (int)(mx >> 23) - 127 + 1;           // floor(log2 max) + 1: max < 2^E
    E = min(max(E, a.e_lo), a.e_hi);
    const float qcap = (float)((1 << m) - 1);
    uint64_t lo = 0, hi = 0;                                         // 16 fields x w bits <= 128 bits
    unsigned bad = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      // scalbnf: |v| * 2^(m-E) without forming 2^(m-E) (it overflows for the smallest exponents)
      const float mag = mx ? fminf(rintf(scalbnf(fabsf(v[j]), m - E)), qcap) : 0.f;
      const uint32_t sgn = (__float_as_uint(v[j]) >> 31) & (mag != 0.f ? 1u : 0u);
      const float back = scalbnf(sgn ? -mag : mag, E - m);
      bad += (back != v[j]) ? 1u : 0u;
      const uint64_t f = (uint64_t)(((uint32_t)mag) | (sgn << m));
      const int bit = j * a.w;
      if (bit < 64) {
        lo |= f << bit;
        if (bit + a.w > 64) hi |= f >> (64 - bit);
      } else {
        hi |= f << (bit - 64);
      }
    }
    if (bad) atomicAdd(a.mismatches, (unsigned long long)bad);
    // destination: group of 256 elements = 16 blocks: [32*w mantissa bytes][16 exponent bytes]
    uint8_t* grp = a.packed + n * a.row_bytes + (b >> 4) * (32 * a.w + 16);
    uint8_t* dm = grp + (b & 15) * 2 * a.w;
    for (int j = 0; j < 2 * a.w; ++j) dm[j] = (uint8_t)((j < 8 ? (lo >> (8 * j)) : (hi >> (8 * (j - 8)))) & 0xffu);
    grp[32 * a.w + (b & 15)] = (uint8_t)(E + a.exp_bias);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
int make_tmap_bf16_kmajor(CUtensorMap* tm, const void* base, int64_t K, int64_t rows, int64_t batch, int64_t ld, int64_t batch_stride,
                          int box_rows);
int make_params(const bq_format* f, FmtParams* p);
int make_tmap_2d(CUtensorMap* tm, const void* base, int dtype_code, int64_t inner, int64_t rows, int64_t row_stride_bytes, int box_inner,
                 int box_rows, int swizzle128);

template <int BN, int XF>
static int launch_xform(const CUtensorMap& tmD, const CUtensorMap& tmR, XArgs g, cudaStream_t st, int kern_id) {
  using Cfg = XCfg<BN, XF>;
  static PerDevice<bool> attr_pd;
  bool& attr = attr_pd.get();
  if (!attr) {
    BQ_CUDA_CHECK(cudaFuncSetAttribute(gemm_xform_kernel<BN, XF>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr = true;
  }
  g.tiles_m = XF == XF_PACKED_A ? 1 : (g.M + kXBM - 1) / kXBM;
  g.tiles_n = XF == XF_PACKED_A ? (g.N + kXBM - 1) / kXBM : (g.N + BN - 1) / BN;
  const int64_t total = (int64_t)g.tiles_m * g.tiles_n;
  if (total > 0x7fffffffll) return BQ_ERR_UNSUPPORTED;
  const int grid = (int)std::min<int64_t>(total, num_sms());
  {
    LaunchScope ls(kern_id, st);
    gemm_xform_kernel<BN, XF><<<grid, kXThreads, Cfg::kSmemBytes, st>>>(tmD, tmR, g);
  }
  BQ_CUDA_CHECK(cudaGetLastError());
  return BQ_OK;
}

static int packed_row_bytes(int w, int64_t K) { return (int)((K / 256) * (32 * w + 16)); }

}  // namespace bq

extern "C" {

// y[M][N] = Q_fx(x)[M][K] @ Wq[N][K]^T (+ bias): the x-quantizer runs in the GEMM prologue (one launch)
int bq_linear_fused(const bq_format* fx, const float* x, int64_t M, int64_t K, int64_t ldx, const void* Wq_bf16, int64_t N,
                    const float* bias, float* y, int64_t ldy, void* stream) {
  using namespace bq;
  if (!fx || M < 0 || N < 0 || K < 0) return BQ_ERR_BAD_ARG;
  if (M == 0 || N == 0) return BQ_OK;
  if (!x || !Wq_bf16 || !y) return BQ_ERR_BAD_ARG;
  if (fx->kind != BQ_KIND_BLOCK_FP && fx->kind != BQ_KIND_BLOCK_MINIFLOAT) return BQ_ERR_UNSUPPORTED;
  if (fx->block_rows != 1 || fx->block_cols != 16) return BQ_ERR_UNSUPPORTED;
  if (K == 0 || (K % 64) || (N % 32)) return BQ_ERR_UNSUPPORTED;
  if ((ldx % 4) || ldx < K || (ldy % 4) || ldy < N || ((uintptr_t)x % 16) || ((uintptr_t)Wq_bf16 % 16) || ((uintptr_t)y % 16) ||
      (bias && ((uintptr_t)bias % 16)))
    return BQ_ERR_BAD_ARG;
  if (M > 0x7fffffff || N > 0x7fffffff || K > 0x7fffffff) return BQ_ERR_UNSUPPORTED;
  XArgs g;
  memset(&g, 0, sizeof(g));
  int rc = make_params(fx, &g.q);
  if (rc) return rc;
  g.q.fold_zero = 0;
  CUtensorMap tmB, tmA;
  if ((rc = make_tmap_bf16_kmajor(&tmB, Wq_bf16, K, N, 1, K, 0, 128))) return rc;
  if ((rc = make_tmap_2d(&tmA, x, /*fp32*/ 0, K, M, ldx * 4, 32, kXBM, 1))) return rc;
  g.C = y; g.bias = bias; g.M = (int)M; g.N = (int)N; g.K = (int)K; g.ldc = ldy;
  return launch_xform<128, XF_QUANT_A>(tmB, tmA, g, (cudaStream_t)stream, kKernGemmXformA);
}

size_t bq_packed_weight_bytes(const bq_format* fw, int64_t N, int64_t K) {
  if (!fw || fw->kind != BQ_KIND_BLOCK_FP || fw->width < 2 || fw->width > 8 || N <= 0 || K <= 0 || (K % 256)) return 0;
  return (size_t)N * (size_t)bq::packed_row_bytes(fw->width, K);
}

int bq_pack_weight(const bq_format* fw, const float* Wq, int64_t N, int64_t K, int64_t ldw, void* packed, unsigned long long* mismatches,
                   void* stream) {
  using namespace bq;
  if (!fw || !Wq || !packed || !mismatches || N <= 0 || K <= 0) return BQ_ERR_BAD_ARG;
  if (fw->kind != BQ_KIND_BLOCK_FP || fw->block_rows != 1 || fw->block_cols != 16) return BQ_ERR_UNSUPPORTED;
  if (fw->width < 2 || fw->width > 8 || fw->exponent_width < 1 || fw->exponent_width > 8 || (K % 256)) return BQ_ERR_UNSUPPORTED;
  if ((ldw % 4) || ldw < K || ((uintptr_t)Wq % 16) || ((uintptr_t)packed % 16)) return BQ_ERR_BAD_ARG;
  PackArgs a;
  a.W = Wq; a.packed = (uint8_t*)packed; a.mismatches = mismatches; a.N = N; a.K = K; a.ldw = ldw;
  a.w = fw->width; a.exp_bias = fw->exponent_bias;
  a.e_lo = -fw->exponent_bias; a.e_hi = (1 << fw->exponent_width) - 1 - fw->exponent_bias;
  a.row_bytes = packed_row_bytes(fw->width, K);
  cudaStream_t st = (cudaStream_t)stream;
  BQ_CUDA_CHECK(cudaMemsetAsync(mismatches, 0, sizeof(unsigned long long), st));
  const int64_t nblk = N * (K / 16);
  {
    LaunchScope ls(kKernPackWeight, st);
    pack_weight_kernel<<<(int)std::min<int64_t>((nblk + 255) / 256, (int64_t)num_sms() * 16), 256, 0, st>>>(a);
  }
  BQ_CUDA_CHECK(cudaGetLastError());
  return BQ_OK;
}

// y[M][N] = A[M][K] (bf16, already x-quantised) @ unpack(packed)[N][K]^T (+ bias): weights decoded in the mainloop
int bq_gemm_packed_tn(const void* A_bf16, const void* packed, const bq_format* fw, float* y, const float* bias, int64_t M, int64_t N,
                      int64_t K, int64_t lda, int64_t ldy, void* stream) {
  using namespace bq;
  if (!fw || M < 0 || N < 0 || K < 0) return BQ_ERR_BAD_ARG;
  if (M == 0 || N == 0) return BQ_OK;
  if (!A_bf16 || !packed || !y) return BQ_ERR_BAD_ARG;
  if (fw->kind != BQ_KIND_BLOCK_FP || fw->width < 2 || fw->width > 8) return BQ_ERR_UNSUPPORTED;
  if (K == 0 || (K % 256) || (N % 32)) return BQ_ERR_UNSUPPORTED;
  if ((lda % 8) || lda < K || (ldy % 4) || ldy < N || ((uintptr_t)A_bf16 % 16) || ((uintptr_t)packed % 16) || ((uintptr_t)y % 16) ||
      (bias && ((uintptr_t)bias % 16)))
    return BQ_ERR_BAD_ARG;
  if (M > 0x7fffffff || N > 0x7fffffff || K > 0x7fffffff) return BQ_ERR_UNSUPPORTED;
  XArgs g;
  memset(&g, 0, sizeof(g));
  g.w = fw->width; g.exp_bias = fw->exponent_bias; g.row_bytes = 32 * fw->width + 16;
  g.C = y; g.bias = bias; g.M = (int)M; g.N = (int)N; g.K = (int)K; g.ldc = ldy;
  CUtensorMap tmA, tmR;
  int rc;
  const int64_t row_bytes_total = packed_row_bytes(fw->width, K);
  if (M <= 128 && (N % 128) == 0) {
    // decode regime: the weight stream is the bound.  Roles swapped — 128 packed weight rows are the M operand, the (few) activation
    // rows the N operand — so that the weights are decoded and read exactly once and the activation re-read is (N/128) * M * K * 2 B
    const int BNx = M <= 32 ? 32 : 128;
    if ((rc = make_tmap_bf16_kmajor(&tmA, A_bf16, K, M, 1, lda, 0, BNx))) return rc;
    if ((rc = make_tmap_2d(&tmR, packed, /*u32*/ 1, row_bytes_total / 4, N, row_bytes_total, g.row_bytes / 4, kXBM, 0))) return rc;
    g.tiles_m = 1;
    if (BNx == 32) return launch_xform<32, XF_PACKED_A>(tmA, tmR, g, (cudaStream_t)stream, kKernGemmXformB);
    return launch_xform<128, XF_PACKED_A>(tmA, tmR, g, (cudaStream_t)stream, kKernGemmXformB);
  }
  const bool small_m = M <= 128 || (N % 128);
  const int BN = small_m ? 32 : 128;
  if ((rc = make_tmap_bf16_kmajor(&tmA, A_bf16, K, M, 1, lda, 0, kXBM))) return rc;
  if ((rc = make_tmap_2d(&tmR, packed, /*u32*/ 1, row_bytes_total / 4, N, row_bytes_total, g.row_bytes / 4, BN, 0))) return rc;
  if (BN == 32) return launch_xform<32, XF_PACKED_B>(tmA, tmR, g, (cudaStream_t)stream, kKernGemmXformB);
  return launch_xform<128, XF_PACKED_B>(tmA, tmR, g, (cudaStream_t)stream, kKernGemmXformB);
}

}  // extern "C"
