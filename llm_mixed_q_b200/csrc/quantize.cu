// quantize.cu — standalone HBM-streaming quantizer kernels (sm_100a) + their C-ABI entry points.
//
// Replaces block()/unblock() (quantizers/utils.py:261-321) and the element math of
// block_fp.py:21-96, block_minifloat.py:22-74, block_log.py:23-69, minifloat.py:21-82,134-196,
// integer.py:25-58 — ~45 ATen launches and >=360 B/element of traffic in the reference — with one
// pass at the algorithmic 8 B/element (fp32 in, fp32 out).
//
// Kernels
//   quant_stream_kernel   the speed path for dense tensors with blocks of 16 along the last dim (and the element-wise kinds):
//                         per-warp bulk-copy (1-D TMA) ring, a lane owns whole blocks, in-place quantise, bulk store.
//   quant_rows_kernel     strided rows / other block widths, and the silu(x) * x2 prologue (Llama down_proj operand).
//                         Blocks of b1 in {4..128} consecutive elements of the
//                         unit-stride last dim (every shipped config: [1,16] / [16]).  One thread
//                         owns 4 consecutive floats (one 16-byte load), a block is b1/4 adjacent
//                         lanes, the shared exponent comes from a warp-shuffle max over those lanes.
//                         Persistent grid (multiple of the SM count), 4 independent 16-byte loads
//                         in flight per thread, streaming cache hints, no shared memory.
//   blocklog_fixup_kernel block_log only: the reference replaces the max of all-zero blocks by the
//                         tensor-wide smallest non-zero block max (block_log.py:50-53), which changes
//                         what zeros quantise to.  The main pass records a per-warp zero-block mask and
//                         the global min; this pass fills the flagged blocks.
//   generic_*             any block shape (2-D blocks, whole-row blocks, odd sizes), any input strides
//                         (k^T views), optional transposed output.  Two passes over a per-block max
//                         workspace.  Correctness path, not a speed path.
//   quant_tile_kernel     k^T views (blocks along a strided dim) through 32x32 shared-memory tiles.
//   norm_quant_warp_kernel / norm_quant_kernel   LayerNorm / RMSNorm fused with the x-quantizers of the consuming Linears.
//   rope_quant_q/k_kernel Llama rotary embedding fused with the two operand quantizers of matmul_0.
//   split3_kernel, split2_f16_rows_kernel        operand planes of the fp32-equivalent split GEMMs.
#include "bq_internal.h"
#include "bq_numerics.cuh"
#include "bq_blockops.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace bq {

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void stg_stream4(float* p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void stg_stream2(void* p, uint32_t a, uint32_t b) {
  asm volatile("st.global.cs.v2.b32 [%0], {%1,%2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t absbits(float v) { return __float_as_uint(v) & 0x7fffffffu; }

template <typename OutT> __device__ __forceinline__ void store4(OutT* y, float4 o);
template <> __device__ __forceinline__ void store4<float>(float* y, float4 o) { stg_stream4(y, o); }
template <> __device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* y, float4 o) {
  stg_stream2(y, pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
}
template <typename OutT> __device__ __forceinline__ void store1(OutT* y, float o);
template <> __device__ __forceinline__ void store1<float>(float* y, float o) { *y = o; }
template <> __device__ __forceinline__ void store1<__nv_bfloat16>(__nv_bfloat16* y, float o) { *y = __float2bfloat16_rn(o); }

// Layout of a "rows" launch.  Index space: every row is padded to slots_per_row 4-float slots
// (a multiple of lanes-per-block) so that a block never straddles a warp.  total_slots < 2^32.
struct RowsGeom {
  uint64_t total_slots;     // n_rows * slots_per_row
  uint32_t slots_per_row;
  uint32_t div_mul, div_shr;  // slot / slots_per_row: t = umulhi(slot, mul); (t + ((slot - t) >> 1)) >> (shr - 1)
  int32_t C;                // valid columns per row (multiple of 4)
  int64_t ldx, ldy;         // row strides in elements
  int32_t lpb;              // lanes per block = b1 / 4 (1 for the element-wise kinds)
  int32_t flat;             // 1: rows are dense and C == padded C -> slot s is elements [4s, 4s+4)
};

constexpr int kThreads = 256;
constexpr int kUnroll = 4;

template <bool FLAT>
__device__ __forceinline__ bool slot_addr(const RowsGeom& g, uint32_t slot, uint32_t total, int64_t& xoff, int64_t& yoff) {
  if (slot >= total) return false;
  if (FLAT) {
    xoff = yoff = (int64_t)slot * 4;
    return true;
  }
  const uint32_t t = __umulhi(slot, g.div_mul);
  const uint32_t row = (t + ((slot - t) >> 1)) >> (g.div_shr - 1);
  const int col = (int)((slot - row * g.slots_per_row) * 4);
  if (col >= g.C) return false;
  xoff = (int64_t)row * g.ldx + col;
  yoff = (int64_t)row * g.ldy + col;
  return true;
}

// Quantise the 4 elements a thread owns, given the block max (as bits).  The literal path is kept out of line:
// it only runs for blocks whose exponents leave the normal range or that contain inf/NaN.
template <int KIND>
__device__ __noinline__ float4 quant4_literal(float4 v, uint32_t mbits, const FmtParams& p) {
  BlockState st;
  st.a = st.b = st.c = 0.f;
  if (IsBlocked<KIND>::value) st = block_state<KIND>(__uint_as_float(mbits), p);
  return make_float4(quant_elem<KIND>(v.x, st, p), quant_elem<KIND>(v.y, st, p), quant_elem<KIND>(v.z, st, p),
                     quant_elem<KIND>(v.w, st, p));
}
// fast arithmetic with per-element cliff checks, out of line (redo path of the streaming kernel; the block's state is valid)
template <int KIND>
__device__ __noinline__ float4 quant4_checked(float4 v, uint32_t mbits, const FmtParams& p) {
  const FastState fs = fast_state<KIND>(mbits, p);
  return make_float4(quant_elem_fast<KIND>(v.x, fs, p), quant_elem_fast<KIND>(v.y, fs, p), quant_elem_fast<KIND>(v.z, fs, p),
                     quant_elem_fast<KIND>(v.w, fs, p));
}
template <int KIND>
__device__ __forceinline__ float4 quant4(float4 v, uint32_t mbits, const FmtParams& p) {
  if (KIND == kInteger || KIND == kNone) {
    BlockState st;
    st.a = st.b = st.c = 0.f;
    return make_float4(quant_elem<KIND>(v.x, st, p), quant_elem<KIND>(v.y, st, p), quant_elem<KIND>(v.z, st, p),
                       quant_elem<KIND>(v.w, st, p));
  }
  const FastState fs = fast_state<KIND>(mbits, p);
  bool ok = fs.ok;
  if (!IsBlocked<KIND>::value) {
    // element-wise kinds: every element must be finite for the fast path
    const uint32_t worst = max(max(absbits(v.x), absbits(v.y)), max(absbits(v.z), absbits(v.w)));
    ok = ok && worst < 0x7f800000u;
  }
  if (!ok) return quant4_literal<KIND>(v, mbits, p);
  return make_float4(quant_elem_fast<KIND>(v.x, fs, p), quant_elem_fast<KIND>(v.y, fs, p), quant_elem_fast<KIND>(v.z, fs, p),
                     quant_elem_fast<KIND>(v.w, fs, p));
}

// gstate[0]: min over non-zero block maxima (uint bits, init 0xffffffff); gstate[1]: 0xffffffff until a zero block is seen
// PRE: the element fed to the quantizer is silu(x) * x2 (silu_mul1, bq_blockops.cuh)
template <int KIND, typename OutT, bool FLAT, bool PRE = false>
__global__ void __launch_bounds__(kThreads) quant_rows_kernel(const float* __restrict__ x, OutT* __restrict__ y, RowsGeom g,
                                                               FmtParams p, uint32_t* __restrict__ gstate,
                                                               uint32_t* __restrict__ zmask,
                                                               const float* __restrict__ x2 = nullptr) {
  constexpr bool kBlocked = IsBlocked<KIND>::value;
  constexpr uint32_t tile = kThreads * kUnroll;
  const uint32_t total = (uint32_t)g.total_slots;
  uint32_t run_min = 0xffffffffu;
  bool saw_zero = false;
  for (uint64_t base64 = (uint64_t)blockIdx.x * tile; base64 < g.total_slots; base64 += (uint64_t)gridDim.x * tile) {
    const uint32_t base = (uint32_t)base64;
    float4 v[kUnroll];
    int64_t yoff[kUnroll];
    bool act[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const uint32_t slot = base + u * kThreads + threadIdx.x;
      int64_t xo = 0;
      yoff[u] = 0;
      act[u] = (base64 + u * kThreads + threadIdx.x < g.total_slots) && slot_addr<FLAT>(g, slot, total, xo, yoff[u]);
      v[u] = act[u] ? ldg_stream4(x + xo) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (PRE) {
        const float4 w = act[u] ? ldg_stream4(x2 + xo) : make_float4(0.f, 0.f, 0.f, 0.f);
        v[u] = make_float4(silu_mul1(v[u].x, w.x), silu_mul1(v[u].y, w.y), silu_mul1(v[u].z, w.z), silu_mul1(v[u].w, w.w));
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      bool zero_block = false;
      uint32_t m = 0x3f800000u;
      if (kBlocked) {
        m = max(max(absbits(v[u].x), absbits(v[u].y)), max(absbits(v[u].z), absbits(v[u].w)));
        for (int o = 1; o < g.lpb; o <<= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
        zero_block = (m == 0);
        if (KIND == kBlockLog) {
          // uniform branch per warp-slot: record which lanes sit in all-zero blocks
          uint32_t zb = __ballot_sync(0xffffffffu, zero_block);
          const uint64_t slot0 = base64 + (uint64_t)u * kThreads + (threadIdx.x & ~31u);
          if ((threadIdx.x & 31) == 0 && slot0 < g.total_slots) zmask[slot0 >> 5] = zb;
          if (act[u]) {
            if (zero_block) saw_zero = true; else run_min = min(run_min, m);
          }
        }
        // block_fp / block_minifloat: an all-zero block only holds pass-through elements, whose result
        // (+0) does not depend on the substituted max (block_fp.py:54-58) — use 1.
        if (zero_block) m = 0x3f800000u;
      }
      if (act[u] && !(KIND == kBlockLog && zero_block)) {
        store4<OutT>(y + yoff[u], quant4<KIND>(v[u], m, p));
      }
    }
  }
  if (KIND == kBlockLog) {
    for (int o = 16; o > 0; o >>= 1) run_min = min(run_min, __shfl_xor_sync(0xffffffffu, run_min, o));
    if ((threadIdx.x & 31) == 0 && run_min != 0xffffffffu) atomicMin(&gstate[0], run_min);
    if (saw_zero) gstate[1] = 0u;
  }
}

template <typename OutT>
__global__ void __launch_bounds__(kThreads) blocklog_fixup_kernel(OutT* __restrict__ y, RowsGeom g, FmtParams p,
                                                                   const uint32_t* __restrict__ gstate,
                                                                   const uint32_t* __restrict__ zmask) {
  if (gstate[1] != 0u) return;   // no all-zero block anywhere
  uint32_t gm = gstate[0];
  // all maxima zero -> ones (block_log.py:50-51); else zero maxima := global min non-zero max (:52-53)
  BlockState st = block_state<kBlockLog>(gm == 0xffffffffu ? 1.0f : __uint_as_float(gm), p);
  float fill = quant_elem<kBlockLog>(0.f, st, p);
  float4 o = make_float4(fill, fill, fill, fill);
  const uint64_t stride = (uint64_t)gridDim.x * kThreads;
  const uint32_t total = (uint32_t)g.total_slots;
  for (uint64_t slot = (uint64_t)blockIdx.x * kThreads + threadIdx.x; slot < g.total_slots; slot += stride) {
    uint32_t zb = zmask[slot >> 5];
    if (!((zb >> (threadIdx.x & 31)) & 1u)) continue;
    int64_t xo, yo;
    if (slot_addr<false>(g, (uint32_t)slot, total, xo, yo)) store4<OutT>(y + yo, o);
  }
}

// ------------------------------------------------------------------------------------------------
// quant_stream_kernel — the speed path: dense tensors, blocks of 16 along the last dim (every shipped config) and the
// element-wise kinds.  Measured motivation (profiles/r01_ncu_quant_rows_s5.json): quant_rows_kernel spends 27 (block_fp) to
// 55 (block_log) instructions per element, ~20 of them on per-slot addressing, predicates, the shuffle reduction and a block
// state that four lanes each recompute.  Here
//   * every WARP is its own pipeline: 4 KB tiles (64 blocks of 16) arrive by one bulk async copy (cp.async.bulk, the 1-D TMA
//     path) into a 3-deep shared-memory ring guarded by mbarriers, are quantised IN PLACE and leave by one bulk store — no
//     per-thread global addressing, no CTA-wide barrier;
//   * a lane owns whole blocks (2 per tile): the block max is 8 integer max instructions on its own registers — no shuffles —
//     and the block state is computed once per 16 elements;
//   * shared-memory reads are 16-byte accesses at a 64-byte lane stride, made conflict-free by rotating the chunk order with
//     the lane index (the element math is order-agnostic).
// ------------------------------------------------------------------------------------------------
constexpr int kStWarps = 8;
#ifndef BQ_ST_NB
#define BQ_ST_NB 2
#endif
#ifndef BQ_ST_STAGES
#define BQ_ST_STAGES 3
#endif
constexpr int kStNB = BQ_ST_NB;                        // blocks per lane per tile
constexpr int kStTileElems = 32 * kStNB * 16;          // 1024 floats
constexpr int kStTileBytes = kStTileElems * 4;
constexpr int kStStages = BQ_ST_STAGES;
constexpr size_t kStSmem = (size_t)kStWarps * kStStages * kStTileBytes + (size_t)kStWarps * kStStages * 8;

__device__ __forceinline__ void st_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void st_bulk_store(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void st_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, P;\n\t}"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity), "r"(0x989680u)
                 : "memory");
  } while (!ok);
}
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_f32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void sts128u(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ uint32_t absmax4(float4 v) {
  return max(max(absbits(v.x), absbits(v.y)), max(absbits(v.z), absbits(v.w)));
}
// one bit per block -> the 4-bits-per-block (one per 4-float slot) layout of zmask that blocklog_fixup_kernel reads
__device__ __forceinline__ uint32_t spread_nibbles(uint32_t b8) {
  uint32_t v = b8 & 0xffu;
  v = (v | (v << 12)) & 0x000f000fu;
  v = (v | (v << 6)) & 0x03030303u;
  v = (v | (v << 3)) & 0x11111111u;
  return v * 0xfu;
}

template <int KIND, typename OutT>
__global__ void __launch_bounds__(kStWarps * 32) quant_stream_kernel(const float* __restrict__ x, OutT* __restrict__ y,
                                                                      uint64_t n_elems, FmtParams p,
                                                                      uint32_t* __restrict__ gstate,
                                                                      uint32_t* __restrict__ zmask) {
  extern __shared__ __align__(128) uint8_t st_smem[];
  constexpr bool kBlocked = IsBlocked<KIND>::value;
  constexpr bool kBf16 = sizeof(OutT) == 2;
  constexpr uint32_t kStageBytes = kStTileBytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t buf0 = (uint32_t)__cvta_generic_to_shared(st_smem) + (uint32_t)warp * kStStages * kStageBytes;
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(st_smem) + (uint32_t)kStWarps * kStStages * kStageBytes +
                        (uint32_t)warp * kStStages * 8;
  auto load_tile = [&](uint32_t dst, uint64_t t, uint32_t bar) {
    const uint64_t left = n_elems - t * kStTileElems;
    st_bulk_load(dst, x + t * kStTileElems, (uint32_t)(left < kStTileElems ? left : kStTileElems) * 4u, bar);
  };
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < kStStages; ++k) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8 * k));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const uint64_t n_tiles = (n_elems + kStTileElems - 1) / kStTileElems;
  const uint64_t n_words = (n_elems + 127) >> 7;                     // zmask words (32 slots of 4 floats each)
  const uint64_t nw = (uint64_t)gridDim.x * kStWarps;
  uint64_t tile = (uint64_t)blockIdx.x * kStWarps + warp;
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < kStStages - 1; ++k) {
      const uint64_t t = tile + (uint64_t)k * nw;
      if (t < n_tiles) load_tile(buf0 + k * kStageBytes, t, bar0 + 8 * k);
    }
  }
  const uint32_t rot = (uint32_t)(lane >> 1) & 3u;
  uint32_t run_min = 0xffffffffu;
  bool saw_zero = false;
  int s = 0;
  uint32_t parity = 0;
  for (; tile < n_tiles; tile += nw) {
    const uint64_t left = n_elems - tile * kStTileElems;
    const uint32_t n_here = (uint32_t)(left < kStTileElems ? left : kStTileElems);
    const uint32_t buf = buf0 + s * kStageBytes;
    st_mbar_wait(bar0 + 8 * s, parity);
    float4 v[kStNB][4];
#pragma unroll
    for (int j = 0; j < kStNB; ++j) {
      const uint32_t base = buf + (uint32_t)(j * 32 + lane) * 64u;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        v[j][c] = lds128(base + (((uint32_t)c + rot) & 3u) * 16u);
      }
    }
    if (kBf16) __syncwarp();                        // every lane holds its inputs before the packed in-place writes below
#pragma unroll
    for (int j = 0; j < kStNB; ++j) {
      const uint32_t blk = (uint32_t)(j * 32 + lane);
      const bool act = blk * 16u < n_here;
      uint32_t m = max(max(absmax4(v[j][0]), absmax4(v[j][1])), max(absmax4(v[j][2]), absmax4(v[j][3])));
      bool zero_block = false;
      bool ok = true;
      if (kBlocked) {
        zero_block = act && (m == 0);
        if (KIND == kBlockLog) {
          const uint32_t zb = __ballot_sync(0xffffffffu, zero_block);
          const uint64_t word = tile * (kStTileElems / 128) + (uint64_t)(j * 4 + lane);
          if (lane < 4 && word < n_words) zmask[word] = spread_nibbles(zb >> (8 * lane));
          if (act) {
            if (zero_block) saw_zero = true; else run_min = min(run_min, m);
          }
        }
        if (m == 0) m = 0x3f800000u;
      } else {
        ok = m < 0x7f800000u;                        // element-wise kinds: all 16 finite
        m = 0x3f800000u;
      }
      if (act) {
        if (KIND == kInteger || KIND == kNone) {
#pragma unroll
          for (int c = 0; c < 4; ++c) v[j][c] = quant4<KIND>(v[j][c], m, p);
        } else {
          const FastState fs = fast_state<KIND>(m, p);
          if (ok && fs.ok) {
            // straight-line shortcut arithmetic for all 16 elements; one test per block for "some element sat within kZone ulps
            // of a log2 rounding cliff", in which case the block is redone through the per-element checked path (out of line)
            uint32_t zacc = 0xffffffffu;
            float4 q[4];
#pragma unroll
            for (int c = 0; c < 4; ++c)
              q[c] = make_float4(quant_elem_fast_impl<KIND, true>(v[j][c].x, fs, p, zacc), quant_elem_fast_impl<KIND, true>(v[j][c].y, fs, p, zacc),
                                 quant_elem_fast_impl<KIND, true>(v[j][c].z, fs, p, zacc), quant_elem_fast_impl<KIND, true>(v[j][c].w, fs, p, zacc));
            if (zone_hit<KIND>(zacc)) {
#pragma unroll
              for (int c = 0; c < 4; ++c) q[c] = quant4_checked<KIND>(v[j][c], m, p);
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) v[j][c] = q[c];
          } else {
#pragma unroll                                       // (a rolled loop would index v dynamically and push it to local memory)
            for (int c = 0; c < 4; ++c) v[j][c] = quant4_literal<KIND>(v[j][c], m, p);
          }
        }
        if (!kBf16) {
          const uint32_t base = buf + blk * 64u;
#pragma unroll
          for (int c = 0; c < 4; ++c) sts128(base + (((uint32_t)c + rot) & 3u) * 16u, v[j][c]);
        } else {
          // chunk c of the lane's rotated order is logical chunk (c + rot) & 3: 8 bytes each, a packed block is 32 bytes
          const uint32_t base = buf + blk * 32u;
#pragma unroll
          for (int c = 0; c < 4; ++c)
            asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(base + (((uint32_t)c + rot) & 3u) * 8u),
                         "r"(pack_bf16x2(v[j][c].x, v[j][c].y)), "r"(pack_bf16x2(v[j][c].z, v[j][c].w))
                         : "memory");
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      st_bulk_store(y + tile * kStTileElems, buf, n_here * (uint32_t)sizeof(OutT));
      // refill the stage whose store was committed one iteration ago (at most this iteration's store may still be reading)
      const uint64_t nt = tile + (uint64_t)(kStStages - 1) * nw;
      if (nt < n_tiles) {
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        const int ns = (s == 0) ? kStStages - 1 : s - 1;
        load_tile(buf0 + ns * kStageBytes, nt, bar0 + 8 * ns);
      }
    }
    if (++s == kStStages) { s = 0; parity ^= 1u; }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  if (KIND == kBlockLog) {
    for (int o = 16; o > 0; o >>= 1) run_min = min(run_min, __shfl_xor_sync(0xffffffffu, run_min, o));
    if (lane == 0 && run_min != 0xffffffffu) atomicMin(&gstate[0], run_min);
    if (saw_zero) gstate[1] = 0u;
  }
}

// ------------------------------------------------------------------------------------------------
// generic path
// ------------------------------------------------------------------------------------------------
struct GenGeom {
  int64_t L, R, C, sL, sR, sC;
  int64_t b0, b1, nb0, nb1;
  int32_t transpose_out;
};
__device__ __forceinline__ void gen_index(const GenGeom& g, int64_t idx, int64_t& l, int64_t& r, int64_t& c) {
  c = idx % g.C;
  int64_t t = idx / g.C;
  r = t % g.R;
  l = t / g.R;
}
__global__ void generic_blockmax_kernel(const float* __restrict__ x, GenGeom g, uint32_t* __restrict__ blkmax) {
  int64_t n = g.L * g.R * g.C;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t l, r, c;
    gen_index(g, idx, l, r, c);
    uint32_t a = absbits(x[l * g.sL + r * g.sR + c * g.sC]);
    if (a) atomicMax(&blkmax[(l * g.nb0 + r / g.b0) * g.nb1 + c / g.b1], a);
  }
}
__global__ void generic_gmin_kernel(const uint32_t* __restrict__ blkmax, int64_t nblk, uint32_t* __restrict__ gstate) {
  uint32_t m = 0xffffffffu;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nblk; i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t v = blkmax[i];
    if (v) m = min(m, v);
  }
  for (int o = 16; o > 0; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m != 0xffffffffu) atomicMin(&gstate[0], m);
}
template <int KIND, typename OutT>
__global__ void generic_quant_kernel(const float* __restrict__ x, OutT* __restrict__ y, GenGeom g, FmtParams p,
                                     const uint32_t* __restrict__ blkmax, const uint32_t* __restrict__ gstate) {
  constexpr bool kBlocked = IsBlocked<KIND>::value;
  int64_t n = g.L * g.R * g.C;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t l, r, c;
    gen_index(g, idx, l, r, c);
    float v = x[l * g.sL + r * g.sR + c * g.sC];
    BlockState st;
    st.a = st.b = st.c = 0.f;
    if (kBlocked) {
      uint32_t m = blkmax[(l * g.nb0 + r / g.b0) * g.nb1 + c / g.b1];
      if (m == 0) {
        uint32_t gm = gstate[0];
        m = (gm == 0xffffffffu) ? __float_as_uint(1.0f) : gm;
      }
      st = block_state<KIND>(__uint_as_float(m), p);
    }
    float o = quant_elem<KIND>(v, st, p);
    int64_t yo = g.transpose_out ? (l * g.C + c) * g.R + r : idx;
    store1<OutT>(y + yo, o);
  }
}


// ------------------------------------------------------------------------------------------------
// tile path: blocks of b1 | 64 along the logical last dim C, input contiguous along C *or* along R
// (k^T views: quantized_functions/matmul.py:187-193 receives key_states.transpose(1,2)), output either
// [L,R,C] or transposed [L,C,R] (K-major hand-off of y to the GEMM).  A 64x64 tile is staged in shared
// memory (pitch 65: conflict-free for both access directions), one thread quantises one block in place,
// and the tile is written back along whichever output dim is contiguous.  block_fp / block_minifloat only
// (all-zero blocks need no tensor-global information there).
// ------------------------------------------------------------------------------------------------
constexpr int kTile = 64;
template <int KIND, typename OutT>
__global__ void __launch_bounds__(256) quant_tile_kernel(const float* __restrict__ x, OutT* __restrict__ y, GenGeom g,
                                                         FmtParams p, int tiles_r, int tiles_c) {
  __shared__ float tile[kTile][kTile + 1];
  const int64_t tiles_per_l = (int64_t)tiles_r * tiles_c;
  const int64_t total = tiles_per_l * g.L;
  const bool in_r_contig = (g.sR == 1 && g.sC != 1);
  const int b1 = (int)g.b1;
  for (int64_t t = blockIdx.x; t < total; t += gridDim.x) {
    const int64_t l = t / tiles_per_l;
    const int64_t rem = t - l * tiles_per_l;
    const int64_t r0 = (rem / tiles_c) * kTile, c0 = (rem % tiles_c) * kTile;
    const float* xb = x + l * g.sL;
    // ---- load (coalesced along the contiguous input dim)
    for (int i = threadIdx.x; i < kTile * kTile; i += 256) {
      int a = i / kTile, b = i % kTile;            // b runs along the contiguous dim
      int rr = in_r_contig ? b : a, cc = in_r_contig ? a : b;
      int64_t r = r0 + rr, c = c0 + cc;
      tile[rr][cc] = (r < g.R && c < g.C) ? xb[r * g.sR + c * g.sC] : 0.f;
    }
    __syncthreads();
    // ---- quantise: one thread per block
    const int blocks_per_row = kTile / b1;
    for (int i = threadIdx.x; i < kTile * blocks_per_row; i += 256) {
      int rr = i % kTile, cb = i / kTile;
      float* bp = &tile[rr][cb * b1];
      uint32_t m = 0;
      for (int j = 0; j < b1; ++j) m = max(m, absbits(bp[j]));
      BlockState st = block_state<KIND>(m == 0 ? 1.0f : __uint_as_float(m), p);
      for (int j = 0; j < b1; ++j) bp[j] = quant_elem<KIND>(bp[j], st, p);
    }
    __syncthreads();
    // ---- store (coalesced along the contiguous output dim)
    for (int i = threadIdx.x; i < kTile * kTile; i += 256) {
      int a = i / kTile, b = i % kTile;
      int rr = g.transpose_out ? b : a, cc = g.transpose_out ? a : b;
      int64_t r = r0 + rr, c = c0 + cc;
      if (r < g.R && c < g.C) {
        int64_t yo = g.transpose_out ? (l * g.C + c) * g.R + r : (l * g.R + r) * g.C + c;
        store1<OutT>(y + yo, tile[rr][cc]);
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static inline float p2(int e) { return ldexpf(1.0f, e); }

int make_params(const bq_format* f, FmtParams* p) {
  memset(p, 0, sizeof(*p));
  p->kind = f->kind;
  p->fold_zero = f->fold_zero ? 1 : 0;
  int mb = 0;
  switch (f->kind) {
    case kBlockFP:
      mb = f->width - 1;
      if (f->exponent_width < 1 || f->exponent_width > 30) return BQ_ERR_BAD_FORMAT;
      p->emin = (float)(-(int64_t)f->exponent_bias);
      p->emax = (float)(((int64_t)1 << f->exponent_width) - 1 - f->exponent_bias);
      break;
    case kBlockMinifloat:
      mb = f->width - f->exponent_width - 1;
      if (f->exponent_width < 1 || f->exponent_width > 30 || f->exponent_bias_width < 1 || f->exponent_bias_width > 30)
        return BQ_ERR_BAD_FORMAT;
      p->bias_hi = (float)(((int64_t)1 << f->exponent_bias_width) - 1);
      p->eb_top = (float)(((int64_t)1 << f->exponent_width) - 1);
      break;
    case kBlockLog:
      if (f->width < 2 || f->width > 31 || f->exponent_bias_width < 1 || f->exponent_bias_width > 30) return BQ_ERR_BAD_FORMAT;
      p->bias_hi = (float)(((int64_t)1 << f->exponent_bias_width) - 1);
      p->eb_top = (float)(((int64_t)1 << (f->width - 1)) - 1);
      break;
    case kMinifloatDenorm:
    case kMinifloatIEEE:
      mb = f->width - f->exponent_width - 1;
      if (f->exponent_width < 1 || f->exponent_width > 30) return BQ_ERR_BAD_FORMAT;
      p->emin = (float)(-(int64_t)f->exponent_bias);
      p->emax = (float)(((int64_t)1 << f->exponent_width) - 1 - f->exponent_bias);
      break;
    case kInteger:
      if (f->width < 1 || f->width > 31) return BQ_ERR_BAD_FORMAT;
      mb = f->exponent_bias;   // frac_width
      if (mb < -60 || mb > 60) return BQ_ERR_BAD_FORMAT;
      p->emin = -(float)((int64_t)1 << (f->width - 1));
      p->emax = (float)(((int64_t)1 << (f->width - 1)) - 1);
      p->shift = p2(mb);
      p->inv_shift = p2(-mb);
      return BQ_OK;
    case kNone:
      return BQ_OK;
    default:
      return BQ_ERR_BAD_FORMAT;
  }
  if (f->kind != kBlockLog) {
    if (mb < 0 || mb > 30) return BQ_ERR_BAD_FORMAT;
    p->shift = p2(mb);
    p->inv_shift = p2(-mb);
    p->qmax = (float)(((int64_t)1 << mb) - 1);
    p->mbits = mb;
  }
  // static part of the fast-path validity (per-block part: fast_state()).  Everything must stay a finite
  // normal number: mantissa <= 22 bits, scalar exponent ranges inside [-100, 100].
  switch (f->kind) {
    case kBlockFP: p->fast_fmt = (mb <= 22 && p->emin >= -1e6f && p->emax <= 1e6f); break;
    case kBlockMinifloat: p->fast_fmt = (mb <= 20 && p->eb_top <= 1e6f && p->bias_hi <= 1e6f); break;
    case kBlockLog: p->fast_fmt = (p->eb_top <= 1e6f && p->bias_hi <= 1e6f); break;
    case kMinifloatDenorm: p->fast_fmt = (mb <= 22 && p->emin >= -100.f && p->emax <= 100.f && p->emax >= p->emin); break;
    case kMinifloatIEEE: p->fast_fmt = (mb <= 20 && p->emin >= -100.f && p->emax <= 100.f && p->emax >= p->emin); break;
    default: p->fast_fmt = 0;
  }
  auto to_i = [](float v) { return (int)fmaxf(fminf(v, 2.0e9f), -2.0e9f); };
  p->emin_i = to_i(p->emin); p->emax_i = to_i(p->emax); p->bias_hi_i = to_i(p->bias_hi); p->eb_top_i = to_i(p->eb_top);
  return BQ_OK;
}

static bool is_blocked(int kind) { return kind == kBlockFP || kind == kBlockMinifloat || kind == kBlockLog; }

// Normalised geometry shared by the workspace query and the launch.
struct Plan {
  bool fast;
  bool tile;
  RowsGeom rg;
  GenGeom gg;
  int64_t nblk;
  size_t ws_bytes;
};

static int make_plan(const bq_format* f, const bq_tensor3* t, int transpose_out, const void* x, const void* y, int y_dtype,
                     Plan* pl) {
  if (t->L < 0 || t->R < 0 || t->C < 0) return BQ_ERR_BAD_ARG;
  const bool blocked = is_blocked(f->kind);
  int64_t b0 = 1, b1 = 1;
  if (blocked) {
    b0 = f->block_rows;
    b1 = f->block_cols;
    if (b0 < 1 || b1 < 1) return BQ_ERR_BAD_ARG;
    if (b0 > t->R && t->R > 0) b0 = t->R;
    if (b1 > t->C && t->C > 0) b1 = t->C;
  }
  GenGeom& gg = pl->gg;
  gg.L = t->L; gg.R = t->R; gg.C = t->C; gg.sL = t->sL; gg.sR = t->sR; gg.sC = t->sC;
  gg.b0 = b0; gg.b1 = b1;
  gg.nb0 = (t->R + b0 - 1) / b0;
  gg.nb1 = (t->C + b1 - 1) / b1;
  gg.transpose_out = transpose_out ? 1 : 0;
  pl->nblk = blocked ? t->L * gg.nb0 * gg.nb1 : 0;

  // fast path eligibility
  bool fast = !transpose_out && t->sC == 1 && (t->C % 4 == 0) && t->C > 0;
  if (blocked) fast = fast && b0 == 1 && b1 >= 4 && b1 <= 128 && (b1 & (b1 - 1)) == 0;
  // rows must collapse to a single row stride
  int64_t n_rows = t->L * t->R;
  int64_t ldx = t->sR;
  if (t->R == 1) ldx = t->sL;
  else if (t->L != 1 && t->sL != t->R * t->sR) fast = false;
  if (n_rows <= 1) ldx = t->C;
  fast = fast && (ldx % 4 == 0) && ldx >= t->C;
  if (x) fast = fast && ((uintptr_t)x % 16 == 0);
  if (y) fast = fast && ((uintptr_t)y % (y_dtype == BQ_F32 ? 16 : 8) == 0);
  pl->fast = fast;
  RowsGeom& rg = pl->rg;
  memset(&rg, 0, sizeof(rg));
  if (fast) {
    int lpb = blocked ? (int)(b1 / 4) : 1;
    int64_t padC = blocked ? gg.nb1 * b1 : t->C;
    rg.slots_per_row = (uint32_t)(padC / 4);
    rg.C = (int32_t)t->C;
    rg.ldx = ldx;
    rg.ldy = t->C;
    rg.lpb = lpb;
    rg.flat = (padC == t->C && ldx == t->C) ? 1 : 0;
    rg.total_slots = (uint64_t)n_rows * rg.slots_per_row;
    if (padC / 4 > 0x7fffffffll || t->C > 0x7fffffffll || rg.total_slots >= 0xffffffffull) pl->fast = false;
    // exact 32-bit division by the invariant d = slots_per_row (libdivide-style, 33-bit magic):
    //   t = umulhi(n, mul);  q = (t + ((n - t) >> 1)) >> (shr - 1)      for every 32-bit n, d >= 2
    // (powers of two: mul = 0 gives q = n >> log2 d).  Verified against n / d in tools/ (see DESIGN.md).
    {
      const uint64_t d = rg.slots_per_row;
      if (d < 2) {
        if (!rg.flat) pl->fast = false;
        rg.div_mul = 0; rg.div_shr = 1;
      } else {
        uint32_t L = 0;
        while ((2ull << L) <= d) ++L;                 // floor(log2 d)
        if ((d & (d - 1)) == 0) { rg.div_mul = 0; rg.div_shr = L; }
        else { rg.div_mul = (uint32_t)(((1ull << (33 + L)) / d + 1) - (1ull << 32)); rg.div_shr = L + 1; }
      }
    }
  }
  pl->tile = !pl->fast && (f->kind == kBlockFP || f->kind == kBlockMinifloat) && b0 == 1 && b1 <= kTile &&
             (b1 & (b1 - 1)) == 0 && (t->sC == 1 || t->sR == 1) && t->C > 0 && t->R > 0;
  size_t ws = 16;
  if (pl->fast) {
    if (f->kind == kBlockLog) ws += 4 * (size_t)((rg.total_slots + 31) / 32);
  } else if (blocked) {
    ws += 4 * (size_t)pl->nblk;
  }
  pl->ws_bytes = ws;
  return BQ_OK;
}

static PerDevice<int> g_num_sms_pd;
static int g_stream_enabled = 1;
void set_stream_quantizer(int on) { g_stream_enabled = on ? 1 : 0; }
int num_sms() {
  int& g_num_sms = g_num_sms_pd.get();
  if (g_num_sms == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
    g_num_sms = n;
  }
  return g_num_sms;
}

template <int KIND, typename OutT>
static int launch_kind(const Plan& pl, const FmtParams& p, const float* x, OutT* y, uint32_t* ws, cudaStream_t st,
                       const float* x2 = nullptr) {
  constexpr bool kBlocked = IsBlocked<KIND>::value;
  uint32_t* gstate = ws;
  uint32_t* aux = ws + 4;
  const int sms = num_sms();
  if (x2 && !(pl.fast && kBlocked)) return BQ_ERR_UNSUPPORTED;      // silu*mul prologue: streaming rows layout, block formats
  if (pl.fast) {
    if (pl.rg.total_slots == 0) return BQ_OK;
    if (KIND == kBlockLog) BQ_CUDA_CHECK(cudaMemsetAsync(gstate, 0xff, 8, st));
    const uint64_t tile = (uint64_t)kThreads * kUnroll;
    uint64_t tiles = (pl.rg.total_slots + tile - 1) / tile;
    // persistent grid: exactly the number of CTAs that are co-resident (one wave), so the grid-stride loop balances
    static PerDevice<int> occ_flat_pd, occ_rows_pd;
    int& occ = pl.rg.flat ? occ_flat_pd.get() : occ_rows_pd.get();
    if (occ == 0) {
      int o = 0;
      cudaError_t e = pl.rg.flat
          ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, quant_rows_kernel<KIND, OutT, true>, kThreads, 0)
          : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, quant_rows_kernel<KIND, OutT, false>, kThreads, 0);
      occ = (e == cudaSuccess && o > 0) ? o : 4;
    }
    int grid = (int)std::min<uint64_t>(tiles, (uint64_t)sms * occ);
    const bool stream = g_stream_enabled && !x2 && pl.rg.flat && (!kBlocked || pl.rg.lpb == 4) && ((uintptr_t)x % 16 == 0) &&
                        ((uintptr_t)y % 16 == 0);
    if (x2) {
      if constexpr (kBlocked) {
        LaunchScope ls(kKernSiluMulQuant, st);
        if (pl.rg.flat) quant_rows_kernel<KIND, OutT, true, true><<<grid, kThreads, 0, st>>>(x, y, pl.rg, p, gstate, aux, x2);
        else quant_rows_kernel<KIND, OutT, false, true><<<grid, kThreads, 0, st>>>(x, y, pl.rg, p, gstate, aux, x2);
      }
    } else if (stream) {
      static PerDevice<bool> attr_pd;
      static PerDevice<int> occ_st_pd;
      bool& attr_set = attr_pd.get();
      int& occ_st = occ_st_pd.get();
      if (!attr_set) {
        BQ_CUDA_CHECK(cudaFuncSetAttribute(quant_stream_kernel<KIND, OutT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStSmem));
        int o = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, quant_stream_kernel<KIND, OutT>, kStWarps * 32, kStSmem);
        occ_st = (e == cudaSuccess && o > 0) ? o : 1;
        attr_set = true;
      }
      const uint64_t n_elems = pl.rg.total_slots * 4;
      const uint64_t wt = (n_elems + kStTileElems - 1) / kStTileElems;
      const int g = (int)std::min<uint64_t>((wt + kStWarps - 1) / kStWarps, (uint64_t)sms * occ_st);
      LaunchScope ls(kKernQuantStream, st);
      quant_stream_kernel<KIND, OutT><<<g, kStWarps * 32, kStSmem, st>>>(x, y, n_elems, p, gstate, aux);
    } else {
      LaunchScope ls(kKernQuantRows, st);
      if (pl.rg.flat) quant_rows_kernel<KIND, OutT, true><<<grid, kThreads, 0, st>>>(x, y, pl.rg, p, gstate, aux);
      else quant_rows_kernel<KIND, OutT, false><<<grid, kThreads, 0, st>>>(x, y, pl.rg, p, gstate, aux);
    }
    if (KIND == kBlockLog) {
      uint64_t blocks = (pl.rg.total_slots + kThreads - 1) / kThreads;
      int g2 = (int)std::min<uint64_t>(blocks, (uint64_t)sms * 8);
      LaunchScope ls(kKernBlockLogFixup, st);
      blocklog_fixup_kernel<OutT><<<g2, kThreads, 0, st>>>(y, pl.rg, p, gstate, aux);
    }
  } else if (pl.tile && (KIND == kBlockFP || KIND == kBlockMinifloat)) {
    int tiles_r = (int)((pl.gg.R + kTile - 1) / kTile), tiles_c = (int)((pl.gg.C + kTile - 1) / kTile);
    int64_t total = (int64_t)tiles_r * tiles_c * pl.gg.L;
    if (total == 0) return BQ_OK;
    int grid = (int)std::min<int64_t>(total, (int64_t)sms * 8);
    LaunchScope ls(kKernQuantTile, st);
    quant_tile_kernel<KIND, OutT><<<grid, 256, 0, st>>>(x, y, pl.gg, p, tiles_r, tiles_c);
  } else {
    int64_t n = pl.gg.L * pl.gg.R * pl.gg.C;
    if (n == 0) return BQ_OK;
    int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)sms * 16);
    if (kBlocked) {
      BQ_CUDA_CHECK(cudaMemsetAsync(gstate, 0xff, 8, st));
      BQ_CUDA_CHECK(cudaMemsetAsync(aux, 0, 4 * (size_t)pl.nblk, st));
      { LaunchScope ls(kKernGenericMax, st); generic_blockmax_kernel<<<grid, 256, 0, st>>>(x, pl.gg, aux); }
      int g2 = (int)std::min<int64_t>((pl.nblk + 255) / 256, (int64_t)sms * 4);
      { LaunchScope ls(kKernGenericMin, st); generic_gmin_kernel<<<g2, 256, 0, st>>>(aux, pl.nblk, gstate); }
    }
    { LaunchScope ls(kKernGenericQuant, st); generic_quant_kernel<KIND, OutT><<<grid, 256, 0, st>>>(x, y, pl.gg, p, aux, gstate); }
  }
  BQ_CUDA_CHECK(cudaGetLastError());
  return BQ_OK;
}

template <typename OutT>
static int launch_dtype(const Plan& pl, const FmtParams& p, const float* x, OutT* y, uint32_t* ws, cudaStream_t st,
                        const float* x2 = nullptr) {
  switch (p.kind) {
    case kBlockFP: return launch_kind<kBlockFP, OutT>(pl, p, x, y, ws, st, x2);
    case kBlockMinifloat: return launch_kind<kBlockMinifloat, OutT>(pl, p, x, y, ws, st, x2);
    case kBlockLog: return launch_kind<kBlockLog, OutT>(pl, p, x, y, ws, st, x2);
    case kMinifloatDenorm: return launch_kind<kMinifloatDenorm, OutT>(pl, p, x, y, ws, st);
    case kMinifloatIEEE: return launch_kind<kMinifloatIEEE, OutT>(pl, p, x, y, ws, st);
    case kInteger: return launch_kind<kInteger, OutT>(pl, p, x, y, ws, st);
    case kNone: return launch_kind<kNone, OutT>(pl, p, x, y, ws, st);
  }
  return BQ_ERR_BAD_FORMAT;
}

int quantize_impl(const bq_format* fmt, const bq_tensor3* t, const float* x, void* y, int y_dtype, int transpose_out,
                  void* ws, size_t ws_bytes, cudaStream_t st, const float* x2) {
  if (!fmt || !t) return BQ_ERR_BAD_ARG;
  FmtParams p;
  int rc = make_params(fmt, &p);
  if (rc) return rc;
  if (y_dtype != BQ_F32 && y_dtype != BQ_BF16) return BQ_ERR_BAD_ARG;
  int64_t n = t->L * t->R * t->C;
  if (n == 0) return BQ_OK;
  if (!x || !y) return BQ_ERR_BAD_ARG;
  if (((uintptr_t)x % 4) || ((uintptr_t)y % (y_dtype == BQ_F32 ? 4 : 2))) return BQ_ERR_BAD_ARG;
  Plan pl;
  rc = make_plan(fmt, t, transpose_out, x, y, y_dtype, &pl);
  if (rc) return rc;
  if (pl.ws_bytes > 16 || is_blocked(fmt->kind)) {
    if (!ws || ws_bytes < pl.ws_bytes) return BQ_ERR_WORKSPACE;
    if ((uintptr_t)ws % 16) return BQ_ERR_BAD_ARG;
  }
  if (x2 && (((uintptr_t)x2 % 16) || !is_blocked(fmt->kind))) return BQ_ERR_UNSUPPORTED;
  if (y_dtype == BQ_F32) return launch_dtype<float>(pl, p, x, (float*)y, (uint32_t*)ws, st, x2);
  return launch_dtype<__nv_bfloat16>(pl, p, x, (__nv_bfloat16*)y, (uint32_t*)ws, st, x2);
}

size_t quantize_ws_bytes(const bq_format* fmt, const bq_tensor3* t) {
  if (!fmt || !t) return 0;
  // worst case over both paths (pointer alignment is unknown here)
  Plan a;
  if (make_plan(fmt, t, 0, nullptr, nullptr, BQ_F32, &a)) return 0;
  size_t w = a.ws_bytes;
  if (is_blocked(fmt->kind)) w = std::max(w, (size_t)16 + 4 * (size_t)a.nblk);
  return (w + 255) & ~(size_t)255;
}

// ------------------------------------------------------------------------------------------------
// fp32 -> three bf16 planes with x == p0 + p1 + p2 up to 2^-25 |x| (error-free splitting by repeated rounding):
// the operand format of the split-precision GEMM that stands in for the reference's UNQUANTISED fp32 matmuls
// (lm_head, modeling_opt.py:942-944; the y operand of block_log matmuls, matmul.py:293-296).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) split3_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int64_t n4,
                                                      int64_t plane) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = ldg_stream4(x + 4 * i);
    float a[4] = {v.x, v.y, v.z, v.w};
    uint32_t p0[2], p1[2], p2[2];
    float h0[4], h1[4], h2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      h0[j] = __bfloat162float(__float2bfloat16_rn(a[j]));
      const float r1 = __fsub_rn(a[j], h0[j]);
      h1[j] = __bfloat162float(__float2bfloat16_rn(r1));
      h2[j] = __bfloat162float(__float2bfloat16_rn(__fsub_rn(r1, h1[j])));
    }
    p0[0] = pack_bf16x2(h0[0], h0[1]); p0[1] = pack_bf16x2(h0[2], h0[3]);
    p1[0] = pack_bf16x2(h1[0], h1[1]); p1[1] = pack_bf16x2(h1[2], h1[3]);
    p2[0] = pack_bf16x2(h2[0], h2[1]); p2[1] = pack_bf16x2(h2[2], h2[3]);
    stg_stream2(out + 4 * i, p0[0], p0[1]);
    stg_stream2(out + plane + 4 * i, p1[0], p1[1]);
    stg_stream2(out + 2 * plane + 4 * i, p2[0], p2[1]);
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm / RMSNorm + quantize: y_k = Q_k(norm(x)) for up to three operand formats, bf16 out.
// Replaces `self_attn_layer_norm` / `final_layer_norm` (models/opt_quantized/modeling_opt.py:386,:414;
// input_layernorm / post_attention_layernorm, models/llama_quantized/modeling_llama.py:386,:399) FOLLOWED BY the
// x-quantizers of the Linears that read the normalised tensor (q/k/v_proj share one input, modeling_opt.py:206,224-225;
// fc1; gate/up_proj) — one read of x, one bf16 write per distinct format, instead of an fp32 round trip per consumer.
// Normalisation arithmetic follows torch's CUDA kernels: LayerNorm  y = fma(gamma, rstd * (x - mean), beta)  with
// rstd = rsqrtf(var + eps) (layer_norm_kernel.cu); RMSNorm  y = w * (x * rsqrtf(mean(x^2) + eps)).  The statistics are
// summed in a different order than torch's Welford / reduction kernels (<= 1-2 ulp in mean / rstd), see DESIGN.md.
//
// Layout (v5): one ROW per CTA at a time, one BLOCK of 16 per thread (H/16 threads, rounded up to whole warps).
//   * rows arrive by one bulk async copy (cp.async.bulk, the 1-D TMA path) into a two-slot shared-memory ring: the next row is
//     in flight while this one is processed; several CTAs per SM cover each other's barriers;
//   * a thread reads its 16 floats ONCE (four 16-byte accesses, chunk order rotated with the lane so that the 64-byte lane
//     stride is conflict-free) and keeps them in registers through all three phases; gamma / beta of the thread's block are
//     loaded once per kernel and live in registers too;
//   * the block max is taken on the thread's own registers and the block state is computed once per block (no shuffles);
//   * the bf16 row is assembled in shared memory and leaves by one bulk store.
// v4 (one row per warp, three rolled passes over shared memory, block = 4 lanes) spent 34 instructions per element and kept 12
// warps per SM (4 at H = 4096): 2.0 TB/s at H = 2048, 1.2 TB/s at H = 4096 (profiles/r01_ncu_micro_s7.json).
// ------------------------------------------------------------------------------------------------
constexpr int kLnMaxThreads = 512;
constexpr int kLnMaxWarps = kLnMaxThreads / 32;
constexpr int kLnMaxStages = 4;          // input ring of 2..4 row slots (LnArgs::stages): the host takes the depth that keeps the most CTAs resident
struct LnArgs {
  const float* x;
  int64_t ldx;
  const float* gamma;
  const float* beta;      // nullptr: RMSNorm
  float eps;
  int H;
  int rows;
  int n_out;
  int stages;
  __nv_bfloat16* out[3];
  FmtParams f[3];
};
__device__ __forceinline__ void bulk_load_row(uint32_t dst, const float* src, uint32_t bytes, uint32_t bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ float warp_sum(float v) {
  for (int o = 16; o > 0; o >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float cta_sum(float v, float* red, int warp, int lane, int nwarp) {
  v = warp_sum(v);
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = red[0];
  for (int w = 1; w < nwarp; ++w) t = __fadd_rn(t, red[w]);
  return t;
}
__global__ void __maxnreg__(80) norm_quant_kernel(LnArgs a) {
  extern __shared__ __align__(128) uint8_t ln_smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarp = blockDim.x >> 5;
  const uint32_t row_bytes = (uint32_t)a.H * 4u, out_bytes = (uint32_t)a.H * 2u;
  const uint32_t in0 = (uint32_t)__cvta_generic_to_shared(ln_smem);
  const int stages = a.stages;
  const uint32_t outb = in0 + (uint32_t)stages * row_bytes;
  float* red = reinterpret_cast<float*>(ln_smem + (size_t)stages * row_bytes + out_bytes);      // [2][kLnMaxWarps]
  const uint32_t bar0 = outb + out_bytes + 2u * kLnMaxWarps * 4u;
  const uint32_t gam0 = bar0 + 8u * kLnMaxStages;         // 16-byte aligned: gamma [H], then beta [H]
  const int stride = gridDim.x;
  int row = blockIdx.x;
  if (tid == 0) {
#pragma unroll
    for (int k = 0; k < kLnMaxStages; ++k) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8u * k));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int k = 0; k < stages - 1; ++k) {
      const int64_t r = (int64_t)row + (int64_t)k * stride;
      if (r < a.rows) bulk_load_row(in0 + (uint32_t)k * row_bytes, a.x + r * a.ldx, row_bytes, bar0 + 8u * k);
    }
  }
  __syncthreads();
  const bool act = tid * 16 < a.H;                       // H % 16 == 0: a thread's block is whole or absent
  const bool ln = a.beta != nullptr;
  // chunk c of the thread's registers holds logical chunk (c + rot) & 3 of its block: conflict-free for the 16-byte loads at a
  // 64-byte lane stride (quarter-warps) AND for the 8-byte packed stores at a 32-byte lane stride (half-warps)
  const uint32_t rot = (uint32_t)((lane >> 1) + (lane >> 3)) & 3u;
  // gamma / beta of the thread's block are parked in shared memory (same rotated chunk order; written and read by the same
  // thread, so no barrier) — holding them in registers cost 32 registers and a CTA of occupancy
  const uint32_t myg = gam0 + (uint32_t)tid * 64u, myb = myg + row_bytes;
  if (act) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint32_t cl = ((uint32_t)c + rot) & 3u;
      sts128(myg + cl * 16u, __ldg(reinterpret_cast<const float4*>(a.gamma + tid * 16 + (int)cl * 4)));
      if (ln) sts128(myb + cl * 16u, __ldg(reinterpret_cast<const float4*>(a.beta + tid * 16 + (int)cl * 4)));
    }
  }
  const float invH = 1.0f / (float)a.H;
  int b = 0;
  uint32_t parity = 0;
  for (; row < a.rows; row += stride) {
    // the slot refilled here held the previous iteration's row: every thread copied it to registers before that iteration's barriers
    const int64_t next = (int64_t)row + (int64_t)(stages - 1) * stride;
    if (tid == 0 && next < a.rows) {
      const uint32_t nb = (uint32_t)(b == 0 ? stages - 1 : b - 1);
      bulk_load_row(in0 + nb * row_bytes, a.x + next * a.ldx, row_bytes, bar0 + 8u * nb);
    }
    st_mbar_wait(bar0 + 8u * b, parity);
    float v[16];
    {
      const uint32_t base = in0 + (uint32_t)b * row_bytes + (uint32_t)tid * 64u;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 t = act ? lds128(base + (((uint32_t)c + rot) & 3u) * 16u) : make_float4(0.f, 0.f, 0.f, 0.f);
        v[4 * c] = t.x; v[4 * c + 1] = t.y; v[4 * c + 2] = t.z; v[4 * c + 3] = t.w;
      }
    }
    // the previous row's bulk store must have finished READING the output slot before anyone rewrites it (after the next barrier)
    if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    if (ln) {
      float s[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) s[c] = __fadd_rn(__fadd_rn(v[4 * c], v[4 * c + 1]), __fadd_rn(v[4 * c + 2], v[4 * c + 3]));
      const float mean = __fmul_rn(cta_sum(__fadd_rn(__fadd_rn(s[0], s[1]), __fadd_rn(s[2], s[3])), red, warp, lane, nwarp), invH);
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = __fsub_rn(v[i], mean);
    }
    float rstd;
    {
      float q[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 16; ++i) q[i & 3] = __fmaf_rn(v[i], v[i], q[i & 3]);
      const float qs = act ? __fadd_rn(__fadd_rn(q[0], q[1]), __fadd_rn(q[2], q[3])) : 0.f;
      rstd = rsqrtf(__fadd_rn(__fmul_rn(cta_sum(qs, red + kLnMaxWarps, warp, lane, nwarp), invH), a.eps));
    }
#pragma unroll 1
    for (int k = 0; k < a.n_out; ++k) {
      const FmtParams& p = (k == 0) ? a.f[0] : ((k == 1) ? a.f[1] : a.f[2]);
      __nv_bfloat16* outp = ((k == 0) ? a.out[0] : ((k == 1) ? a.out[1] : a.out[2])) + (int64_t)row * a.H;
      if (k > 0) {
        if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncthreads();
      }
      if (act) {
        float y[16];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t cl = ((uint32_t)c + rot) & 3u;
          const float4 gv = lds128(myg + cl * 16u);
          if (ln) {
            const float4 bv = lds128(myb + cl * 16u);
            y[4 * c] = __fmaf_rn(gv.x, __fmul_rn(rstd, v[4 * c]), bv.x);
            y[4 * c + 1] = __fmaf_rn(gv.y, __fmul_rn(rstd, v[4 * c + 1]), bv.y);
            y[4 * c + 2] = __fmaf_rn(gv.z, __fmul_rn(rstd, v[4 * c + 2]), bv.z);
            y[4 * c + 3] = __fmaf_rn(gv.w, __fmul_rn(rstd, v[4 * c + 3]), bv.w);
          } else {
            y[4 * c] = __fmul_rn(gv.x, __fmul_rn(v[4 * c], rstd));
            y[4 * c + 1] = __fmul_rn(gv.y, __fmul_rn(v[4 * c + 1], rstd));
            y[4 * c + 2] = __fmul_rn(gv.z, __fmul_rn(v[4 * c + 2], rstd));
            y[4 * c + 3] = __fmul_rn(gv.w, __fmul_rn(v[4 * c + 3], rstd));
          }
        }
        quantize_signed16_rt(y, p);
        const uint32_t base = outb + (uint32_t)tid * 32u;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(base + (((uint32_t)c + rot) & 3u) * 8u),
                       "r"(pack_bf16x2(y[4 * c], y[4 * c + 1])), "r"(pack_bf16x2(y[4 * c + 2], y[4 * c + 3]))
                       : "memory");
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (tid == 0) st_bulk_store(outp, outb, out_bytes);
    }
    if (++b == stages) { b = 0; parity ^= 1u; }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// v6, rows of at most 2048 elements, one kind for every output: one ROW per WARP, a lane owns NB whole blocks (blocks lane,
// lane + 32, ...) and holds them in registers.  No CTA barrier and no cross-warp traffic at all: the statistics are two warp
// shuffle trees, the row slot is handed back to the bulk-copy engine as soon as the row is in registers (the next row lands
// while this one is processed), gamma / beta are read from one shared copy per CTA.  ~16 instructions per element against 33
// for v5 (whose per-row CTA barriers, partial-sum exchange and thread-0 bookkeeping are per-row costs paid by 4 warps).
constexpr int kLwWarps = 8;              // rows in flight per CTA
// FULL: the row has exactly NB * 32 * G blocks (H = 512 * NB * G), no per-block predicate anywhere.
// G: warps per row.  G = 1 serves H <= 2048 (a lane holds up to 4 blocks = 64 values); G = 2 serves H <= 4096 (OPT-6.7B, Llama-7B: the
// row-per-CTA kernel ran those at 3.7 TB/s) with the same 64 values per lane: warp gw of a row's pair owns blocks (j * G + gw) * 32 + lane,
// the two statistics are exchanged through 8 bytes of shared memory per warp and one 64-thread named barrier each, and the pair's
// first lane drives the row slot's bulk copies.
template <int KIND, int NB, bool FULL, int G>
__global__ void __launch_bounds__(kLwWarps * G * 32, G == 1 ? 2 : 1) norm_quant_warp_kernel(LnArgs a) {
  extern __shared__ __align__(128) uint8_t ln_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rg = warp / G, gw = warp % G;                // row slot of this warp, position inside the row's warp group
  const uint32_t row_bytes = (uint32_t)a.H * 4u, out_bytes = (uint32_t)a.H * 2u;
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(ln_smem);
  const uint32_t gam0 = sbase, bet0 = sbase + row_bytes;
  const uint32_t in0 = sbase + 2u * row_bytes + (uint32_t)rg * (row_bytes + out_bytes);
  const uint32_t outb = in0 + row_bytes;
  const uint32_t bars = sbase + 2u * row_bytes + (uint32_t)kLwWarps * (row_bytes + out_bytes);
  const uint32_t bar = bars + 8u * (uint32_t)rg;
  const uint32_t xch = bars + 8u * (uint32_t)kLwWarps + (uint32_t)rg * (2u * G * 4u);      // [2 statistics][G warps] partial sums
  const bool ln = a.beta != nullptr;
  const bool leader = lane == 0 && gw == 0;
  const int nblk = a.H >> 4;
  const int stride = gridDim.x * kLwWarps;
  int row = blockIdx.x * kLwWarps + rg;
  if (leader) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // programmatic dependent launch (see sm100_ptx.cuh: griddep_wait): the barrier set-up and the copy of gamma / beta — model
  // parameters, never written by a kernel of the forward — overlap the tail of the previous kernel; the activation rows are read
  // and the outputs written only after that kernel has completed
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  for (int i = threadIdx.x; i < (a.H >> 2); i += kLwWarps * G * 32) {
    sts128(gam0 + 16u * i, __ldg(reinterpret_cast<const float4*>(a.gamma) + i));
    if (ln) sts128(bet0 + 16u * i, __ldg(reinterpret_cast<const float4*>(a.beta) + i));
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (leader && row < a.rows) bulk_load_row(in0, a.x + (int64_t)row * a.ldx, row_bytes, bar);
  __syncthreads();
  const uint32_t rot = (uint32_t)((lane >> 1) + (lane >> 3)) & 3u;
  const float invH = 1.0f / (float)a.H;
  uint32_t parity = 0;
  auto group_bar = [&]() {
    if (G > 1) asm volatile("bar.sync %0, %1;" ::"r"(1 + rg), "r"(32 * G) : "memory");
    else __syncwarp();
  };
  // sum over the row's G warps of a per-warp total; every warp adds the partials in the same order, so all agree bit for bit
  auto group_sum = [&](float v, int which) {
    v = warp_sum(v);
    if (G == 1) return v;
    if (lane == 0) sts_f32(xch + (uint32_t)(which * G + gw) * 4u, v);
    group_bar();
    float t = lds_f32(xch + (uint32_t)(which * G) * 4u);
#pragma unroll
    for (int w = 1; w < G; ++w) t = __fadd_rn(t, lds_f32(xch + (uint32_t)(which * G + w) * 4u));
    return t;
  };
  for (; row < a.rows; row += stride) {
    // (G > 1) the previous row's store must have left the output slot before ANY warp of the group rewrites it: the leader waits
    // here, ahead of the first group barrier of this row
    if (G > 1 && leader) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    st_mbar_wait(bar, parity);
    parity ^= 1u;
    float v[NB][16];
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const int blk = (j * G + gw) * 32 + lane;
      const uint32_t base = in0 + (uint32_t)blk * 64u;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 t = (FULL || blk < nblk) ? lds128(base + (((uint32_t)c + rot) & 3u) * 16u) : make_float4(0.f, 0.f, 0.f, 0.f);
        v[j][4 * c] = t.x; v[j][4 * c + 1] = t.y; v[j][4 * c + 2] = t.z; v[j][4 * c + 3] = t.w;
      }
    }
    // the row is in registers once the group has passed its next barrier: refill the slot while the row is processed
    auto refill = [&]() {
      if (leader) {
        const int64_t next = (int64_t)row + stride;
        if (next < a.rows) bulk_load_row(in0, a.x + next * a.ldx, row_bytes, bar);
        if (G == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // the previous row's store has left the output slot
      }
    };
    if (G == 1) {
      __syncwarp();
      refill();
    }
    float mean = 0.f;
    if (ln) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        float t[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) t[c] = __fadd_rn(__fadd_rn(v[j][4 * c], v[j][4 * c + 1]), __fadd_rn(v[j][4 * c + 2], v[j][4 * c + 3]));
        s = __fadd_rn(s, __fadd_rn(__fadd_rn(t[0], t[1]), __fadd_rn(t[2], t[3])));
      }
      mean = __fmul_rn(group_sum(s, 0), invH);
      if (G > 1) refill();
    }
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      float t[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        v[j][i] = __fsub_rn(v[j][i], mean);
        t[i & 3] = __fmaf_rn(v[j][i], v[j][i], t[i & 3]);
      }
      if (FULL || (j * G + gw) * 32 + lane < nblk) q = __fadd_rn(q, __fadd_rn(__fadd_rn(t[0], t[1]), __fadd_rn(t[2], t[3])));
    }
    const float rstd = rsqrtf(__fadd_rn(__fmul_rn(group_sum(q, 1), invH), a.eps));     // (G = 1: warp_sum's shuffles also order lane 0's wait above)
    if (G > 1 && !ln) refill();
#pragma unroll 1
    for (int k = 0; k < a.n_out; ++k) {
      const FmtParams& p = (k == 0) ? a.f[0] : ((k == 1) ? a.f[1] : a.f[2]);
      __nv_bfloat16* outp = ((k == 0) ? a.out[0] : ((k == 1) ? a.out[1] : a.out[2])) + (int64_t)row * a.H;
      if (k > 0) {
        if (leader) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        group_bar();
      }
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const int blk = (j * G + gw) * 32 + lane;
        if (FULL || blk < nblk) {
          float y[16];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint32_t cl = ((uint32_t)c + rot) & 3u;
            const float4 gv = lds128(gam0 + (uint32_t)blk * 64u + cl * 16u);
            if (ln) {
              const float4 bv = lds128(bet0 + (uint32_t)blk * 64u + cl * 16u);
              y[4 * c] = __fmaf_rn(gv.x, __fmul_rn(rstd, v[j][4 * c]), bv.x);
              y[4 * c + 1] = __fmaf_rn(gv.y, __fmul_rn(rstd, v[j][4 * c + 1]), bv.y);
              y[4 * c + 2] = __fmaf_rn(gv.z, __fmul_rn(rstd, v[j][4 * c + 2]), bv.z);
              y[4 * c + 3] = __fmaf_rn(gv.w, __fmul_rn(rstd, v[j][4 * c + 3]), bv.w);
            } else {
              y[4 * c] = __fmul_rn(gv.x, __fmul_rn(v[j][4 * c], rstd));
              y[4 * c + 1] = __fmul_rn(gv.y, __fmul_rn(v[j][4 * c + 1], rstd));
              y[4 * c + 2] = __fmul_rn(gv.z, __fmul_rn(v[j][4 * c + 2], rstd));
              y[4 * c + 3] = __fmul_rn(gv.w, __fmul_rn(v[j][4 * c + 3], rstd));
            }
          }
          quantize_signed16<KIND>(y, p);
          const uint32_t base = outb + (uint32_t)blk * 32u;
#pragma unroll
          for (int c = 0; c < 4; ++c)
            asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(base + (((uint32_t)c + rot) & 3u) * 8u),
                         "r"(pack_bf16x2(y[4 * c], y[4 * c + 1])), "r"(pack_bf16x2(y[4 * c + 2], y[4 * c + 3]))
                         : "memory");
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      group_bar();
      if (leader) st_bulk_store(outp, outb, out_bytes);
    }
  }
  if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int KIND, int NB, bool FULL, int G>
static int launch_norm_quant_warp(const LnArgs& a, cudaStream_t st) {
  const size_t smem = (size_t)2 * a.H * 4 + (size_t)kLwWarps * ((size_t)a.H * 6) + kLwWarps * 8 + kLwWarps * 2 * G * 4;
  static PerDevice<size_t> smem_attr_pd;
  size_t& smem_attr = smem_attr_pd.get();
  if (smem > smem_attr) {
    BQ_CUDA_CHECK(cudaFuncSetAttribute(norm_quant_warp_kernel<KIND, NB, FULL, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_attr = smem;
  }
  static PerDevice<int> occ_h_pd, occ_pd;
  int& occ_h = occ_h_pd.get();
  int& occ = occ_pd.get();
  if (occ_h != a.H) {
    int o = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, norm_quant_warp_kernel<KIND, NB, FULL, G>, kLwWarps * G * 32, smem) != cudaSuccess || o < 1) o = 1;
    occ = o;
    occ_h = a.H;
  }
  const int64_t want = ((int64_t)a.rows + kLwWarps - 1) / kLwWarps;
  const int grid = (int)std::min<int64_t>(want, (int64_t)num_sms() * occ);
  {
    LaunchScope ls(kKernLnQuant, st);
    BQ_CUDA_CHECK(launch_ex(norm_quant_warp_kernel<KIND, NB, FULL, G>, grid, kLwWarps * G * 32, smem, st, 1, a));
  }
  BQ_CUDA_CHECK(cudaGetLastError());
  return BQ_OK;
}
template <int KIND>
static int launch_norm_quant_warp_nb(const LnArgs& a, cudaStream_t st) {
  const int nblk = a.H / 16;
  if (nblk == 256) return launch_norm_quant_warp<KIND, 4, true, 2>(a, st);
  if (nblk > 128) return launch_norm_quant_warp<KIND, 4, false, 2>(a, st);
  if (nblk == 128) return launch_norm_quant_warp<KIND, 4, true, 1>(a, st);
  if (nblk == 64) return launch_norm_quant_warp<KIND, 2, true, 1>(a, st);
  if (nblk <= 32) return launch_norm_quant_warp<KIND, 1, false, 1>(a, st);
  if (nblk <= 64) return launch_norm_quant_warp<KIND, 2, false, 1>(a, st);
  return launch_norm_quant_warp<KIND, 4, false, 1>(a, st);
}
static bool g_ln_warp_rows = true;
void set_ln_warp_rows(int on) { g_ln_warp_rows = on != 0; }

int norm_quantize_impl(const float* x, int64_t rows, int64_t H, int64_t ldx, const float* gamma, const float* beta, float eps,
                       int n_out, const bq_format* fmts, void* const* outs, cudaStream_t st) {
  if (rows < 0 || H <= 0 || n_out < 1 || n_out > 3 || !fmts || !outs) return BQ_ERR_BAD_ARG;
  if (rows == 0) return BQ_OK;
  if (!x || !gamma) return BQ_ERR_BAD_ARG;
  if ((H % 16) || H > 16 * kLnMaxThreads || rows > 0x7fffffff) return BQ_ERR_UNSUPPORTED;      // one block of 16 per thread
  if (((uintptr_t)x % 16) || (ldx % 4) || ldx < H || ((uintptr_t)gamma % 16) || (beta && ((uintptr_t)beta % 16))) return BQ_ERR_BAD_ARG;
  LnArgs a;
  memset(&a, 0, sizeof(a));
  a.x = x; a.ldx = ldx; a.gamma = gamma; a.beta = beta; a.eps = eps; a.H = (int)H; a.n_out = n_out;
  for (int k = 0; k < n_out; ++k) {
    if (!outs[k] || ((uintptr_t)outs[k] % 16)) return BQ_ERR_BAD_ARG;
    if (fmts[k].kind != BQ_KIND_BLOCK_FP && fmts[k].kind != BQ_KIND_BLOCK_MINIFLOAT && fmts[k].kind != BQ_KIND_BLOCK_LOG) return BQ_ERR_UNSUPPORTED;
    if (fmts[k].block_rows != 1 || fmts[k].block_cols != 16) return BQ_ERR_UNSUPPORTED;
    int rc = make_params(&fmts[k], &a.f[k]);
    if (rc) return rc;
    a.f[k].fold_zero = 0;
    a.out[k] = (__nv_bfloat16*)outs[k];
  }
  a.rows = (int)rows;
  if (g_ln_warp_rows && H <= 4096) {
    bool same = true;
    for (int k = 1; k < n_out; ++k) same = same && fmts[k].kind == fmts[0].kind;
    if (same) {
      a.stages = 1;
      if (fmts[0].kind == BQ_KIND_BLOCK_LOG) return launch_norm_quant_warp_nb<kBlockLog>(a, st);
      return fmts[0].kind == BQ_KIND_BLOCK_FP ? launch_norm_quant_warp_nb<kBlockFP>(a, st) : launch_norm_quant_warp_nb<kBlockMinifloat>(a, st);
    }
  }
  const int threads = (int)((H / 16 + 31) / 32) * 32;
  auto smem_for = [&](int stages) { return (size_t)stages * H * 4 + (size_t)H * 2 + 2 * kLnMaxWarps * 4 + kLnMaxStages * 8 + (size_t)2 * H * 4; };
  static PerDevice<size_t> smem_attr_pd;
  size_t& smem_attr = smem_attr_pd.get();
  if (smem_for(kLnMaxStages) > smem_attr) {
    BQ_CUDA_CHECK(cudaFuncSetAttribute(norm_quant_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_for(kLnMaxStages)));
    smem_attr = smem_for(kLnMaxStages);
  }
  // ring depth: the deepest of 4 / 3 / 2 slots that does not cost a resident CTA
  static PerDevice<int> occ_cache_h_pd, occ_cache_pd, stages_cache_pd;
  int& occ_cache_h = occ_cache_h_pd.get();
  int& occ_cache = occ_cache_pd.get();
  int& stages_cache = stages_cache_pd.get();
  if (occ_cache_h != (int)H) {
    int best = 0, best_s = 2;
    for (int s = kLnMaxStages; s >= 2; --s) {
      int o = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, norm_quant_kernel, threads, smem_for(s)) != cudaSuccess) o = 0;
      if (o > best) { best = o; best_s = s; }
    }
    occ_cache = std::max(best, 1);
    stages_cache = best_s;
    occ_cache_h = (int)H;
  }
  a.stages = stages_cache;
  const size_t smem = smem_for(a.stages);
  const int grid = (int)std::min<int64_t>(rows, (int64_t)num_sms() * occ_cache);
  {
    LaunchScope ls(kKernLnQuant, st);
    norm_quant_kernel<<<grid, threads, smem, st>>>(a);
  }
  BQ_CUDA_CHECK(cudaGetLastError());
  return BQ_OK;
}

// ------------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------------
// Rotary position embedding + the x / y quantizers of matmul_0, token-major (Llama).  Replaces apply_rotary_pos_emb
// (quantized_functions/rotary_positional_encoding.py:142-167: q_embed = (q * cos) + (rotate_half(q) * sin) with the already
// quantised cos / sin tables gathered by position_ids) FOLLOWED BY the two operand quantizers of the QK^T matmul
// (models/llama_quantized/modeling_llama.py:309-314, quantized_functions/matmul.py:166-196): q in blocks of 16 along d,
// k^T in blocks of 16 consecutive key positions at a fixed feature.  The reference runs ~12 element-wise kernels (two gathers, four
// multiplies, two negations, two concatenations, two additions) and two quantizer calls over [B, S, H] fp32 tensors; here q and k
// are read once and the bf16 operands of the attention kernel are written once.  Same arithmetic, same order: rn(rn(x * cos) +
// rn(rot * sin)), rot = -x[i + d/2] for i < d/2 and x[i - d/2] otherwise (negation commutes with the rounding of the product).
// ------------------------------------------------------------------------------------------------
struct RopeArgs {
  const float* q;
  const float* k;
  const float* cs;            // cos table [rows][d]
  const float* sn;            // sin table [rows][d]
  const int64_t* pos;         // [B][S] or nullptr (position = s)
  int64_t table_rows;         // rows of cs / sn: explicit positions are clamped to [0, table_rows) — never an out-of-bounds read;
                              // range errors are the host mirror's to raise (IndexError, like the reference's cos[position_ids])
  __nv_bfloat16* Qq;
  __nv_bfloat16* Kq;
  int B, S, heads, d;
  int64_t ldq, ldk;           // token strides of q / k (elements); outputs are dense [B*S][heads*d]
  FmtParams fq, fk;
};
// q: one thread per block of 16 along d
__global__ void __launch_bounds__(256) rope_quant_q_kernel(RopeArgs a) {
  const int H = a.heads * a.d, bpt = H >> 4, half = a.d >> 1;
  const int64_t nblk = (int64_t)a.B * a.S * bpt;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nblk; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t tok = i / bpt;
    const int f0 = (int)(i - tok * bpt) << 4;              // first feature of the block
    const int e0 = f0 % a.d;                               // position inside the head
    const bool lo = e0 < half;
    const int64_t p = a.pos ? min(max(a.pos[tok], (int64_t)0), a.table_rows - 1) : (tok % a.S);
    const float* x = a.q + tok * a.ldq + f0;
    const float* xp = x + (lo ? half : -half);
    const float* c = a.cs + p * a.d + e0;
    const float* s = a.sn + p * a.d + e0;
    float y[16];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 xv = *reinterpret_cast<const float4*>(x + 4 * j), pv = *reinterpret_cast<const float4*>(xp + 4 * j);
      const float4 cv = __ldg(reinterpret_cast<const float4*>(c + 4 * j)), sv = __ldg(reinterpret_cast<const float4*>(s + 4 * j));
      const float r0 = lo ? -pv.x : pv.x, r1 = lo ? -pv.y : pv.y, r2 = lo ? -pv.z : pv.z, r3 = lo ? -pv.w : pv.w;
      y[4 * j] = __fadd_rn(__fmul_rn(xv.x, cv.x), __fmul_rn(r0, sv.x));
      y[4 * j + 1] = __fadd_rn(__fmul_rn(xv.y, cv.y), __fmul_rn(r1, sv.y));
      y[4 * j + 2] = __fadd_rn(__fmul_rn(xv.z, cv.z), __fmul_rn(r2, sv.z));
      y[4 * j + 3] = __fadd_rn(__fmul_rn(xv.w, cv.w), __fmul_rn(r3, sv.w));
    }
    quantize_signed16_rt(y, a.fq);
    uint4* o = reinterpret_cast<uint4*>(a.Qq + tok * H + f0);
    o[0] = make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
    o[1] = make_uint4(pack_bf16x2(y[8], y[9]), pack_bf16x2(y[10], y[11]), pack_bf16x2(y[12], y[13]), pack_bf16x2(y[14], y[15]));
  }
}
// k: one thread per (16 consecutive key positions, feature); a warp covers 32 consecutive features, so every access is a
// coalesced 128-byte (fp32) or 64-byte (bf16) row segment
__global__ void __launch_bounds__(256) rope_quant_k_kernel(RopeArgs a) {
  const int H = a.heads * a.d, half = a.d >> 1, sblk = a.S >> 4;
  const int64_t n = (int64_t)a.B * sblk * H;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int f = (int)(i % H);
    const int64_t bs = i / H;                              // (batch, block of 16 positions)
    const int b = (int)(bs / sblk), s0 = (int)(bs - (int64_t)b * sblk) << 4;
    const int e = f % a.d;
    const bool lo = e < half;
    const int fp = lo ? f + half : f - half;
    float y[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) {
      const int64_t tok = (int64_t)b * a.S + s0 + t;
      const int64_t p = a.pos ? min(max(a.pos[tok], (int64_t)0), a.table_rows - 1) : (int64_t)(s0 + t);
      const float xv = a.k[tok * a.ldk + f], pv = a.k[tok * a.ldk + fp];
      const float cv = __ldg(a.cs + p * a.d + e), sv = __ldg(a.sn + p * a.d + e);
      y[t] = __fadd_rn(__fmul_rn(xv, cv), __fmul_rn(lo ? -pv : pv, sv));
    }
    quantize_signed16_rt(y, a.fk);
#pragma unroll
    for (int t = 0; t < 16; ++t) a.Kq[((int64_t)b * a.S + s0 + t) * H + f] = __float2bfloat16_rn(y[t]);
  }
}

int rope_quantize_impl(const float* q, const float* k, const float* cos_t, const float* sin_t, const int64_t* pos, int64_t table_rows,
                       int64_t B, int64_t S, int heads, int d, int64_t ldq, int64_t ldk, const bq_format* fq, const bq_format* fk,
                       void* Qq, void* Kq, cudaStream_t st) {
  if (B < 0 || S < 0 || heads <= 0 || d <= 0 || !fq || !fk) return BQ_ERR_BAD_ARG;
  if (B == 0 || S == 0) return BQ_OK;
  if (!q || !k || !cos_t || !sin_t || !Qq || !Kq) return BQ_ERR_BAD_ARG;
  if ((d % 32) || (S % 16)) return BQ_ERR_UNSUPPORTED;              // a block of 16 stays inside one half of a head / inside the sequence
  if (table_rows < 1 || (!pos && table_rows < S)) return BQ_ERR_BAD_ARG;
  const int64_t H = (int64_t)heads * d;
  if (ldq < H || ldk < H || (ldq % 4) || (ldk % 4) || ((uintptr_t)q % 16) || ((uintptr_t)k % 16) || ((uintptr_t)cos_t % 16) ||
      ((uintptr_t)sin_t % 16) || ((uintptr_t)Qq % 16) || ((uintptr_t)Kq % 2))
    return BQ_ERR_BAD_ARG;
  if (B * S * H > 0x7fffffffffffll || H > 0x7fffffff) return BQ_ERR_UNSUPPORTED;
  RopeArgs a;
  memset(&a, 0, sizeof(a));
  for (int i = 0; i < 2; ++i) {
    const bq_format* f = i ? fk : fq;
    if (f->kind != BQ_KIND_BLOCK_FP && f->kind != BQ_KIND_BLOCK_MINIFLOAT) return BQ_ERR_UNSUPPORTED;
    if (f->block_rows != 1 || f->block_cols != 16) return BQ_ERR_UNSUPPORTED;
    int rc = make_params(f, i ? &a.fk : &a.fq);
    if (rc) return rc;
    (i ? a.fk : a.fq).fold_zero = 0;
  }
  a.q = q; a.k = k; a.cs = cos_t; a.sn = sin_t; a.pos = pos; a.table_rows = table_rows; a.Qq = (__nv_bfloat16*)Qq; a.Kq = (__nv_bfloat16*)Kq;
  a.B = (int)B; a.S = (int)S; a.heads = heads; a.d = d; a.ldq = ldq; a.ldk = ldk;
  const int64_t nq = B * S * (H / 16), nk = B * (S / 16) * H;
  const int64_t cap = (int64_t)num_sms() * 16;
  {
    LaunchScope ls(kKernRopeQuant, st);
    rope_quant_q_kernel<<<(int)std::min<int64_t>((nq + 255) / 256, cap), 256, 0, st>>>(a);
  }
  BQ_CUDA_CHECK(cudaGetLastError());
  {
    LaunchScope ls(kKernRopeQuant, st);
    rope_quant_k_kernel<<<(int)std::min<int64_t>((nk + 255) / 256, cap), 256, 0, st>>>(a);
  }
  BQ_CUDA_CHECK(cudaGetLastError());
  return BQ_OK;
}

// ------------------------------------------------------------------------------------------------
// fp32 -> two fp16 planes per row with a per-row power-of-two scale:  x * 2^e = hi + lo (+- 2^-22 of the row max),
// hi = fp16(x * 2^e), lo = fp16(x * 2^e - hi), e chosen so that the row max lands in [2^14, 2^15).
// Operand format of the fp16-split GEMM that stands in for the reference's UNQUANTISED fp32 matmuls (lm_head): three
// plane products (lo*hi, hi*lo, hi*hi) reproduce the fp32 product to ~2^-21 relative — below the accumulation-order
// noise of an fp32 GEMM — at half the tensor work of the 6-term bf16 split.  inv_scale[r] = 2^-e undoes the scaling in
// the GEMM epilogue (exact).  One warp per row, two passes (the second re-reads the row from L1/L2).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) split2_f16_rows_kernel(const float* __restrict__ x, int64_t ldx, int rows, int K,
                                                               __half* __restrict__ planes, float* __restrict__ inv_scale) {
  const int lane = threadIdx.x & 31;
  const int nslot = K >> 2;
  const int64_t plane = (int64_t)rows * K;
  for (int row = blockIdx.x * 8 + (threadIdx.x >> 5); row < rows; row += gridDim.x * 8) {
    const float* xr = x + (int64_t)row * ldx;
    uint32_t m = 0;
    for (int s = lane; s < nslot; s += 32) {
      const float4 v = *reinterpret_cast<const float4*>(xr + 4 * s);
      m = max(m, max(max(absbits(v.x), absbits(v.y)), max(absbits(v.z), absbits(v.w))));
    }
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    // exponent of the row max (finite, non-zero rows); scale = 2^(14 - floor(log2 max)), clamped to normal powers of two
    int ex = (int)(m >> 23) - 127;
    if (m == 0 || m >= 0x7f800000u) ex = 14;
    int e = 14 - ex;
    e = min(max(e, -126), 126);
    const float sc = __int_as_float((e + 127) << 23);
    if (lane == 0) inv_scale[row] = __int_as_float((127 - e) << 23);
    __half* h0 = planes + (int64_t)row * K;
    __half* h1 = h0 + plane;
    for (int s = lane; s < nslot; s += 32) {
      const float4 v = *reinterpret_cast<const float4*>(xr + 4 * s);
      const float a[4] = {__fmul_rn(v.x, sc), __fmul_rn(v.y, sc), __fmul_rn(v.z, sc), __fmul_rn(v.w, sc)};
      __half hi[4], lo[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        hi[j] = __float2half_rn(a[j]);
        lo[j] = __float2half_rn(__fsub_rn(a[j], __half2float(hi[j])));
      }
      uint2 ph, pl;
      ph.x = (uint32_t)__half_as_ushort(hi[0]) | ((uint32_t)__half_as_ushort(hi[1]) << 16);
      ph.y = (uint32_t)__half_as_ushort(hi[2]) | ((uint32_t)__half_as_ushort(hi[3]) << 16);
      pl.x = (uint32_t)__half_as_ushort(lo[0]) | ((uint32_t)__half_as_ushort(lo[1]) << 16);
      pl.y = (uint32_t)__half_as_ushort(lo[2]) | ((uint32_t)__half_as_ushort(lo[3]) << 16);
      *reinterpret_cast<uint2*>(h0 + 4 * s) = ph;
      *reinterpret_cast<uint2*>(h1 + 4 * s) = pl;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// exhaustive self-test of the exponent-field shortcuts against libdevice log2f
// ------------------------------------------------------------------------------------------------
__global__ void selftest_log2_kernel(unsigned long long* mism) {
  unsigned long long bad0 = 0, bad1 = 0, bad2 = 0;
  const uint64_t n = 0x7f800000ull;      // every positive finite pattern incl. denormals (0 excluded below)
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x + 1; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const float x = __uint_as_float((uint32_t)i);
    const float l = log2f(x);
    const int c = (int)ceilf(l), f = (int)floorf(l), r = (int)rintf(l);
    bad0 += (ceil_log2_i(x) != c);
    bad1 += (floor_log2_i(x) != f);
    bad2 += (rint_log2_i(x) != r);
    // biased per-element variants: checked == exact; unchecked == exact whenever the zone key says "not near a cliff"
    uint32_t z0 = 0xffffffffu, z1 = 0xffffffffu, z2 = 0xffffffffu, zu = 0xffffffffu;
    if (i >= 0x00800000ull) {
      bad0 += (ceil_log2_biased_nf<false>(x, zu) != c + 127);
      bad1 += (floor_log2_biased_nf<false>(x, zu) != f + 127);
      bad2 += (rint_log2_biased_f<false>(x, zu) != r + 127);
      const int cn = ceil_log2_biased_nf<true>(x, z0), fn = floor_log2_biased_nf<true>(x, z1), rn = rint_log2_biased_f<true>(x, z2);
      bad0 += (!zone_hit<kMinifloatDenorm>(z0) && cn != c + 127);
      bad1 += (!zone_hit<kBlockMinifloat>(z1) && fn != f + 127);
      bad2 += (!zone_hit<kBlockLog>(z2) && rn != r + 127);
    } else {
      bad2 += (rint_log2_biased_f<false>(x, zu) > 1) + (rint_log2_biased_f<true>(x, z2) > 1);   // denormals: anything <= 1
    }
    // the scaled ("deep") variant of the block_log fast path: every pattern below 2^100, denormals included
    if (i < ((100ull + 127ull) << 23)) {
      uint32_t z3 = 0xffffffffu;
      bad2 += (rint_log2_biased_deep<false>(x, zu) != r + 127);
      const int rd = rint_log2_biased_deep<true>(x, z3);
      bad2 += (!zone_hit<kBlockLog>(z3) && rd != r + 127);
    }
  }
  if (bad0) atomicAdd(&mism[0], bad0);
  if (bad1) atomicAdd(&mism[1], bad1);
  if (bad2) atomicAdd(&mism[2], bad2);
}

}  // namespace bq

extern "C" {
int bq_split3_bf16(const float* x, void* planes_bf16, int64_t n, void* stream) {
  if (n < 0) return BQ_ERR_BAD_ARG;
  if (n == 0) return BQ_OK;
  if (!x || !planes_bf16 || (n % 4) || ((uintptr_t)x % 16) || ((uintptr_t)planes_bf16 % 8)) return BQ_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n4 = n / 4;
  int grid = (int)std::min<int64_t>((n4 + 255) / 256, (int64_t)bq::num_sms() * 8);
  {
    bq::LaunchScope ls(bq::kKernSplit3, st);
    bq::split3_kernel<<<grid, 256, 0, st>>>(x, (__nv_bfloat16*)planes_bf16, n4, n);
  }
  BQ_CUDA_CHECK(cudaGetLastError());
  return BQ_OK;
}
int bq_split2_f16_rows(const float* x, int64_t rows, int64_t K, int64_t ldx, void* planes_f16, float* inv_scale, void* stream) {
  if (rows < 0 || K < 0) return BQ_ERR_BAD_ARG;
  if (rows == 0 || K == 0) return BQ_OK;
  if (!x || !planes_f16 || !inv_scale || (K % 4) || (ldx % 4) || ldx < K || ((uintptr_t)x % 16) || ((uintptr_t)planes_f16 % 8))
    return BQ_ERR_BAD_ARG;
  if (rows > 0x7fffffff || K > 0x7fffffff) return BQ_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  int grid = (int)std::min<int64_t>((rows + 7) / 8, (int64_t)bq::num_sms() * 8);
  {
    bq::LaunchScope ls(bq::kKernSplit3, st);
    bq::split2_f16_rows_kernel<<<grid, 256, 0, st>>>(x, ldx, (int)rows, (int)K, (__half*)planes_f16, inv_scale);
  }
  BQ_CUDA_CHECK(cudaGetLastError());
  return BQ_OK;
}
void bq_set_stream_quantizer(int on) { bq::set_stream_quantizer(on); }
void bq_set_norm_warp_rows(int on) { bq::set_ln_warp_rows(on); }
int bq_rope_quantize(const float* q, const float* k, const float* cos_table, const float* sin_table, const int64_t* position_ids,
                     int64_t table_rows, int64_t B, int64_t S, int32_t heads, int32_t head_dim, int64_t ldq, int64_t ldk,
                     const bq_format* fq, const bq_format* fk, void* Qq_bf16, void* Kq_bf16, void* stream) {
  return bq::rope_quantize_impl(q, k, cos_table, sin_table, position_ids, table_rows, B, S, heads, head_dim, ldq, ldk, fq, fk, Qq_bf16,
                                Kq_bf16, (cudaStream_t)stream);
}
int bq_selftest_log2(unsigned long long* mismatches_dev3, void* stream) {
  if (!mismatches_dev3) return BQ_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  BQ_CUDA_CHECK(cudaMemsetAsync(mismatches_dev3, 0, 3 * sizeof(unsigned long long), st));
  bq::selftest_log2_kernel<<<bq::num_sms() * 8, 256, 0, st>>>(mismatches_dev3);
  BQ_CUDA_CHECK(cudaGetLastError());
  return BQ_OK;
}
int bq_norm_quantize(const float* x, int64_t rows, int64_t H, int64_t ldx, const float* gamma, const float* beta, float eps,
                     int32_t n_out, const bq_format* fmts, void* const* outs_bf16, void* stream) {
  return bq::norm_quantize_impl(x, rows, H, ldx, gamma, beta, eps, n_out, fmts, outs_bf16, (cudaStream_t)stream);
}
size_t bq_quantize_workspace_bytes(const bq_format* fmt, const bq_tensor3* x) { return bq::quantize_ws_bytes(fmt, x); }
int bq_quantize(const bq_format* fmt, const bq_tensor3* x_desc, const float* x, void* y, int32_t y_dtype,
                int32_t transpose_out, void* ws, size_t ws_bytes, void* stream) {
  return bq::quantize_impl(fmt, x_desc, x, y, y_dtype, transpose_out, ws, ws_bytes, (cudaStream_t)stream);
}
int bq_silu_mul_quantize(const bq_format* fmt, const bq_tensor3* desc, const float* gate, const float* up, void* y, int32_t y_dtype,
                         void* ws, size_t ws_bytes, void* stream) {
  if (!up) return BQ_ERR_BAD_ARG;
  return bq::quantize_impl(fmt, desc, gate, y, y_dtype, 0, ws, ws_bytes, (cudaStream_t)stream, up);
}
}
