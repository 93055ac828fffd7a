// gemm_sm100.cu — persistent, warp-specialised bf16 GEMM for sm_100a:
//   TMA (cp.async.bulk.tensor, 128B swizzle) -> smem ring -> tcgen05.mma (kind::f16, fp32 accumulate in
//   TMEM, double-buffered) -> tcgen05.ld epilogue (+bias) -> fp32 global.
//
//   C[b][m][n] = sum_k A[b][m][k] * B[b][n][k] (+ bias[n])       both operands K-major ("TN")
//
// It is the multiply of the reference's quantized Linear / matmul / bmm
// (quantized_modules/linear.py:71 F.linear, quantized_functions/matmul.py:196 torch.matmul/bmm, fp32 there):
// block-quantised operands with <= 8 significant bits are exact in bf16 and their products are exact
// in fp32, so only the accumulation order differs from the reference's fp32 GEMM.
//
// Warp roles (256 threads, or 384 with a fused epilogue; 1 CTA per SM, persistent, static round-robin tile schedule):
//   warp 0   TMA producer (one elected lane)      warp 1   MMA issuer (one elected lane)
//   warp 2   TMEM allocator                       warps 4-7 (4-11) epilogue (TMEM lane quarter = warp % 4)
// Pipelines: smem full/empty ring (TMA <-> MMA), TMEM full/empty x2 accumulators (MMA <-> epilogue).
// CG == 2 (the instance every large Linear uses): a cluster of two CTAs shares one 256 x 256 tile through cta_group::2 MMAs —
// each CTA loads half of A and half of B, the leader issues, commits are multicast to both CTAs' barriers.
// Tile raster: N-fastest, or M-fastest inside bands of row blocks when B cannot stay in L2 (GemmArgs::band_m).
// Variants: plain (+bias), fused epilogue (EpiArgs; compile-time specialised instances), split precision (plane pairs accumulated
// into one TMEM accumulator: the fp32-equivalent lm_head and the batched general-route matmul).
#include "bq_internal.h"
#include "bq_blockops.cuh"
#include "sm100_ptx.cuh"

namespace bq {

constexpr int kBM = 128;          // UMMA M (cta_group::1)
constexpr int kBK = 64;           // 64 bf16 = 128 bytes = one SWIZZLE_128B atom row
constexpr int kUmmaK = 16;

constexpr int kEpiWarpWords = 32 * 33 + 64 * 4;      // per epilogue warp: 32x33 transpose tile + 64 uint4 block states
// BN: tile width; EPI: fused epilogue (8 epilogue warps + scratch) ; CG: CTAs per MMA (cta_group::1 / ::2).
// EPI: 0 none; 1 generic (every EpiArgs combination, runtime dispatch); specialised instances of the same code with the mode
// fixed at compile time (each path gets its own register allocation — in the generic instance the fp32 / row-block paths cost the
// bf16 column-block path 7 %): 2 = fp32 out, no quantiser (coalesced store / residual / replicas); 3 = blocks along N, bf16 out;
// 4 = blocks along M, bf16 out (3, 4: no residual, no replicas); 5 = gated SiLU over interleaved gate / up column groups (EpiArgs::act 2);
// 6 = rotary position embedding + block quantiser (either direction), bf16 out (EpiArgs::rope_cos).
// With CG == 2 a CTA pair computes a 256 x 256 tile: each CTA owns 128 rows of A and of the accumulator and HALF of the B tile,
// which the pair's MMA reads from both shared memories — 2/3 of the smem fill traffic and operand reads per FLOP of CG == 1.
template <int BN, int EPI = 0, int CG = 1> struct GemmCfg {
  static constexpr int kEpiWarps = EPI ? 8 : 4;
  static constexpr int kThreads = 128 + 32 * kEpiWarps;
  static constexpr int kStageA = kBM * kBK * 2;
  static constexpr int kStageB = (BN / CG) * kBK * 2;
  static constexpr int kStage = kStageA + kStageB;
  static constexpr int kEpiBytes = EPI ? kEpiWarps * kEpiWarpWords * 4 : 0;
  static constexpr int kBudget = 224 * 1024 - 2048 - kEpiBytes;      // 227 KB per CTA minus barriers / alignment slack / scratch
  static constexpr int kStages = kBudget / kStage > 8 ? 8 : kBudget / kStage;
  static constexpr int kTmemCols = 2 * BN;                     // two accumulators; power of two >= 32
  static constexpr int kBarBytes = (2 * kStages + 4) * 8 + 16;
  static constexpr int kEpiOff = kStages * kStage + ((kBarBytes + 15) / 16) * 16;
  static constexpr int kSmemBytes = kEpiOff + kEpiBytes + 1024;   // + alignment slack
  static_assert(kStages >= 3, "pipeline too shallow");
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
};

// Fused epilogue (EPI kernels): v = acc + bias; v *= scale; v = act(v); v = residual + v; v = Q(v); store fp32 / bf16.
// Replaces the separate ATen / quantizer launches between two quantized Linears of a decoder layer
// (reference models/opt_quantized/modeling_opt.py:206-225 q*scaling, :412-420 fc2(relu(fc1(x))), residual adds :395,:424,
// and the x-quantizer of the consuming op, quantized_modules/linear.py:63-71 / quantized_functions/matmul.py:165-193).
struct EpiArgs {
  const float* residual;   // fp32 [M][ldr] or nullptr
  int64_t ldr;
  float scale;             // 1.0f: skipped
  int act;                 // 0 none, 1 ReLU, 2 gated SiLU: B's rows come in groups of 32 = 16 gate rows + the 16 up rows of the same
                           // features; the epilogue forms silu(gate) * up, block-quantises the 16 results and stores bf16 [M][N / 2]
  int out_bf16;            // 0: fp32 store, 1: bf16 store
  int qmode;               // 0 none; 1: blocks of 16 along N (one thread's registers); 2: blocks of 16 along M (16 lanes)
  FmtParams q;
  // fused all-gather: the tile is also stored to the same position of n_rep peer-mapped copies of C (NVLink stores)
  int n_rep;
  void* rep[BQ_MAX_REPLICAS];
  // rotary position embedding between the accumulator and the quantiser (EPI 6; Llama q_proj / k_proj -> matmul_0 operands):
  // the N columns are heads of rope_d features, row m is the token at position rope_pos[m] (or m % rope_S)
  const float* rope_cos;   // [rope_rows][rope_d] tables, already quantised by the host like the reference's; nullptr: no RoPE
  const float* rope_sin;
  const int64_t* rope_pos; // [M] or nullptr
  int64_t rope_rows;
  int rope_S, rope_d;
  // EPI 6, q | k | v in ONE launch: the N columns are up to three segments of seg_cols columns (the projections' quantised weights
  // concatenated along N, all reading the same x operand); segment s stores bf16 [M][ldc] to seg_C[s] with its own block direction,
  // format and RoPE flag.  seg_cols == 0: one segment described by C / qmode / q above.
  int seg_cols;
  void* seg_C[3];
  int seg_qmode[3];
  int seg_rope[3];
  FmtParams seg_q[3];
};

struct GemmArgs {
  float* C;
  const float* bias;
  int M, N, K, batch;
  int64_t ldc, sc;
  int tiles_m, tiles_n;
  int band_m;        // > 0: tiles are walked M-fastest inside bands of band_m row blocks (B too large to stay in L2), else N-fastest
  int b_broadcast;   // B has no batch dim (weights)
  // split-precision mode (n_terms > 0, batch == 1): A and B are stacks of bf16 planes [planes][rows][K]; the K loop runs
  // over the (plane_a, plane_b) pairs below and accumulates every product into the same TMEM accumulator
  int n_terms;
  int8_t term_a[8], term_b[8];
  int fp16;                    // operands are fp16 (kind::f16 with F16 inputs) instead of bf16
  // causal structure of attention matmuls (square per-batch problems, rows = queries): 1 = output tiles that lie wholly above the
  // diagonal are neither computed nor stored (QK^T: the consumer never reads them); 2 = the K loop stops at the last key a tile's
  // rows can see (P @ V: P is zero behind the diagonal)
  int causal;
  const float* row_scale;      // optional exact power-of-two output scales: C = acc * row_scale[m] * col_scale[n] (+ bias)
  const float* col_scale;
  EpiArgs epi;
};

// Blocks of 16 consecutive ROWS (qmode 2): the 32x32 chunk of |v| bit patterns is transposed through shared memory so that
// lane c reduces column c (two blocks: rows 0-15, 16-31) and evaluates the per-block state ONCE; the states are then
// broadcast-read by the row-owning lanes.  Common case (every block of the chunk on the fast path, one warp vote): a straight-line
// loop — 32 independent state loads and element chains, no per-element branch.  (v1 reduced with 4 shuffles per element and
// re-derived the state in all 16 lanes: 238 TFLOP/s; v2 tested a per-element state flag, which serialised the 32 elements behind
// 32 dependent load -> branch -> math chains: k_proj 152 us against 97 us for q_proj, ncu.)
template <int KIND>
__device__ __noinline__ float quant_with_max_cold(float x, uint32_t mbits, FmtParams p) { return quant_with_max<KIND>(x, mbits, p); }
template <int KIND>
__device__ __forceinline__ void quant_rowblocks32(float (&v)[32], const FmtParams& q, uint32_t* tb, int lane) {
  uint4* st = reinterpret_cast<uint4*>(tb + 32 * 33);
#pragma unroll
  for (int j = 0; j < 32; ++j) tb[lane * 33 + j] = __float_as_uint(v[j]) & 0x7fffffffu;
  __syncwarp();
  uint32_t m0 = 0, m1 = 0;
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    m0 = max(m0, tb[r * 33 + lane]);
    m1 = max(m1, tb[(r + 16) * 33 + lane]);
  }
  uint4 sf[2], ss[2];
  bool fast = true;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const uint32_t mraw = h ? m1 : m0;
    const uint32_t m = mraw ? mraw : 0x3f800000u;                // all-zero block: every element is +-0 and passes through as +0
    const FastState fs = fast_state<KIND>(m, q);
    fast = fast && fs.ok;
    if (KIND == kBlockFP) sf[h] = make_uint4(__float_as_uint(fs.f0), __float_as_uint(fs.f1), __float_as_uint(fs.c0), __float_as_uint(fs.c1));
    else sf[h] = make_uint4((uint32_t)fs.i0, (uint32_t)fs.i1, (uint32_t)fs.i2, 0u);
    ss[h] = make_uint4(m, mraw ? 0u : 2u, 0u, 0u);
  }
  const bool all_fast = __all_sync(0xffffffffu, fast);
  const int half = lane >> 4;
  if (all_fast) {
    st[lane] = sf[0];
    st[32 + lane] = sf[1];
    __syncwarp();
    FastState fs;
    fs.ok = true;
    fs.f0 = fs.f1 = fs.c0 = fs.c1 = 0.f;
    fs.i0 = fs.i1 = fs.i2 = 0;
    fs.hi = __fadd_rn(kRintMagic, q.qmax);
    if (KIND != kBlockFP) { fs.f0 = __fadd_rn(kRintMagic, q.shift); fs.f1 = __fadd_rn(fs.f0, q.qmax); }
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const uint4 s = st[half * 32 + j];
      if (KIND == kBlockFP) {
        fs.f0 = __uint_as_float(s.x); fs.f1 = __uint_as_float(s.y); fs.c0 = __uint_as_float(s.z); fs.c1 = __uint_as_float(s.w);
      } else {
        fs.i0 = (int)s.x; fs.i1 = (int)s.y; fs.i2 = (int)s.z;
      }
      v[j] = quant_elem_fast<KIND>(v[j], fs, q);
    }
  } else {
    st[lane] = ss[0];
    st[32 + lane] = ss[1];
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 32; ++j) {                               // rare: out-of-line per element
      const uint4 s = st[half * 32 + j];
      v[j] = (s.y == 2u) ? 0.f : quant_with_max_cold<KIND>(v[j], s.x, q);
    }
  }
  __syncwarp();                                                  // scratch is reused by the next chunk
}

// 32x32 fp32 chunk, row-per-lane registers -> "coalesced" registers: o[i] = columns 4*(lane&7)..+3 of row 4*i + (lane>>3), through the
// warp's scratch as 16-byte accesses whose chunk position is XOR-swizzled with the row (conflict-free both ways).  A store
// instruction then covers four full 128-byte row segments instead of 32 rows x 16 bytes.
__device__ __forceinline__ void chunk_to_coalesced(const float (&v)[32], float4 (&o)[8], uint32_t* scratch, int lane) {
  float4* sc = reinterpret_cast<float4*>(scratch);
#pragma unroll
  for (int j = 0; j < 8; ++j) sc[lane * 8 + (j ^ (lane & 7))] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  __syncwarp();
  const int rsub = lane >> 3, c = lane & 7;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = i * 4 + rsub;
    o[i] = sc[r * 8 + (c ^ (r & 7))];
  }
  __syncwarp();                                                  // scratch is reused by the next chunk
}

// one 32-column chunk of one accumulator row through the fused epilogue
template <int EM>
__device__ __forceinline__ void epilogue_chunk(const GemmArgs& g, const uint32_t (&r)[32], const float4 (&res)[8], int row, int col0,
                                               bool row_ok, uint32_t* scratch, int lane, float* Cb) {
  const EpiArgs& e = g.epi;
  const int qmode = (EM == 1) ? e.qmode : (EM == 2 ? 0 : (EM == 3 ? 1 : 2));
  const bool out_bf16 = (EM == 1) ? (e.out_bf16 != 0) : (EM != 2);
  const bool has_residual = (EM == 1 || EM == 2) && e.residual != nullptr;
  const int n_rep = (EM == 1 || EM == 2) ? e.n_rep : 0;
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
  if (g.bias) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 bv = *reinterpret_cast<const float4*>(g.bias + col0 + j);
      v[j] = __fadd_rn(v[j], bv.x); v[j + 1] = __fadd_rn(v[j + 1], bv.y);
      v[j + 2] = __fadd_rn(v[j + 2], bv.z); v[j + 3] = __fadd_rn(v[j + 3], bv.w);
    }
  }
  if (EM == 5) {
    // Gated SiLU (Llama MLP, reference models/llama_quantized/modeling_llama.py:84 down_proj(act_fn(gate_proj(x)) * up_proj(x))): this
    // 32-column chunk holds gate[f .. f+16) and up[f .. f+16) of the same 16 features (weights interleaved by the host), i.e. exactly one
    // block of down_proj's x-quantizer in this thread's registers.  Replaces two fp32 GEMM outputs (8 B/elem written, 8 read back) and
    // the silu*mul+quantise kernel by one 2 B/elem store; same silu_mul1 / quantize_signed16 as that kernel, so the same bits.
    if (!row_ok) return;
    float t[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) t[i] = silu_mul1(v[i], v[16 + i]);
    quantize_signed16_rt(t, e.q);
    __nv_bfloat16* c = reinterpret_cast<__nv_bfloat16*>(Cb) + (int64_t)row * g.ldc + (col0 >> 1);
    reinterpret_cast<uint4*>(c)[0] = make_uint4(pack_bf16_rn(t[0], t[1]), pack_bf16_rn(t[2], t[3]), pack_bf16_rn(t[4], t[5]), pack_bf16_rn(t[6], t[7]));
    reinterpret_cast<uint4*>(c)[1] = make_uint4(pack_bf16_rn(t[8], t[9]), pack_bf16_rn(t[10], t[11]), pack_bf16_rn(t[12], t[13]), pack_bf16_rn(t[14], t[15]));
    return;
  }
  if (e.scale != 1.0f) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __fmul_rn(v[j], e.scale);
  }
  if (e.act == 1) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = (v[j] < 0.f) ? 0.f : v[j];      // torch relu: NaN propagates
  }
  if (qmode == 0 && !out_bf16) {                               // warp-uniform
    // fp32 output without a quantiser (out_proj / fc2 residual epilogues, plain Linear, fused all-gather): residual read, local
    // store and the peers' copies all run in the coalesced layout.  (Lane-per-row 16-byte accesses made the K = 2048 residual
    // epilogue longer than its mainloop — 133 us against 97 us — and moved 16-byte packets over NVLink: 163 GB/s.)
    const int rsub = lane >> 3, c4 = (lane & 7) * 4, wrow0 = row - lane;
    float4 rs[8];
    if (has_residual) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rr = wrow0 + i * 4 + rsub;
        rs[i] = rr < g.M ? *reinterpret_cast<const float4*>(e.residual + (int64_t)rr * e.ldr + col0 + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    float4 o[8];
    chunk_to_coalesced(v, o, scratch, lane);
    if (has_residual) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        o[i].x = __fadd_rn(rs[i].x, o[i].x); o[i].y = __fadd_rn(rs[i].y, o[i].y);
        o[i].z = __fadd_rn(rs[i].z, o[i].z); o[i].w = __fadd_rn(rs[i].w, o[i].w);
      }
    }
#pragma unroll 1
    for (int p = -1; p < n_rep; ++p) {                           // -1: this rank's C, then the peers' copies of the gathered output
      float* cp = (p < 0 ? Cb : reinterpret_cast<float*>(e.rep[p])) + col0 + c4;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rr = wrow0 + i * 4 + rsub;
        if (rr < g.M) *reinterpret_cast<float4*>(cp + (int64_t)rr * g.ldc) = o[i];
      }
    }
    return;
  }
  if (has_residual) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[4 * j] = __fadd_rn(res[j].x, v[4 * j]); v[4 * j + 1] = __fadd_rn(res[j].y, v[4 * j + 1]);
      v[4 * j + 2] = __fadd_rn(res[j].z, v[4 * j + 2]); v[4 * j + 3] = __fadd_rn(res[j].w, v[4 * j + 3]);
    }
  }
  if (qmode == 1) {
#pragma unroll
    for (int blk = 0; blk < 2; ++blk) {
      float t[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) t[i] = v[blk * 16 + i];
      quantize_signed16_rt(t, e.q);
#pragma unroll
      for (int i = 0; i < 16; ++i) v[blk * 16 + i] = t[i];
    }
  } else if (qmode == 2) {
    // a block = 16 consecutive rows (lanes 0-15 / 16-31) of one column; M % 16 == 0, so a block is all-valid or all-invalid
    if (e.q.kind == kBlockFP) quant_rowblocks32<kBlockFP>(v, e.q, scratch, lane);
    else quant_rowblocks32<kBlockMinifloat>(v, e.q, scratch, lane);
  }
  if (out_bf16 && n_rep > 0) {
    // fused all-gather of a quantised bf16 operand (tensor-parallel layer: fc1's output in fc2's format, dist.py): the warp's
    // 32 rows x 64 bytes are transposed through its scratch so that one store instruction covers 8 rows x 64 contiguous bytes
    // (lane-per-row 16-byte packets crossed NVLink at a fifth of the link rate); the local copy and the peers' copies use the
    // same registers.  All 32 lanes take part (rows beyond M are skipped at the store).
    uint4* sc = reinterpret_cast<uint4*>(scratch);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      sc[lane * 4 + (j ^ (lane & 3))] = make_uint4(pack_bf16_rn(v[8 * j], v[8 * j + 1]), pack_bf16_rn(v[8 * j + 2], v[8 * j + 3]),
                                                   pack_bf16_rn(v[8 * j + 4], v[8 * j + 5]), pack_bf16_rn(v[8 * j + 6], v[8 * j + 7]));
    __syncwarp();
    const int rsub = lane >> 2, c = lane & 3, wrow0 = row - lane;
    uint4 o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = i * 8 + rsub;
      o[i] = sc[r * 4 + (c ^ (r & 3))];
    }
    __syncwarp();                                                // scratch is reused by the next chunk
#pragma unroll 1
    for (int p = -1; p < n_rep; ++p) {                           // -1: this rank's C, then the peers' copies
      __nv_bfloat16* cp = (p < 0 ? reinterpret_cast<__nv_bfloat16*>(Cb) : reinterpret_cast<__nv_bfloat16*>(e.rep[p])) + col0 + c * 8;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int rr = wrow0 + i * 8 + rsub;
        if (rr < g.M) *reinterpret_cast<uint4*>(cp + (int64_t)rr * g.ldc) = o[i];
      }
    }
    return;
  }
  if (!row_ok) return;
  const int64_t off = (int64_t)row * g.ldc + col0;
  if (out_bf16) {
    uint4 o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      o[j] = make_uint4(pack_bf16_rn(v[8 * j], v[8 * j + 1]), pack_bf16_rn(v[8 * j + 2], v[8 * j + 3]),
                        pack_bf16_rn(v[8 * j + 4], v[8 * j + 5]), pack_bf16_rn(v[8 * j + 6], v[8 * j + 7]));
    __nv_bfloat16* c = reinterpret_cast<__nv_bfloat16*>(Cb) + off;      // (replicas of a bf16 result took the coalesced path above)
#pragma unroll
    for (int j = 0; j < 4; ++j) reinterpret_cast<uint4*>(c)[j] = o[j];
  } else {
    float* c = Cb + off;
#pragma unroll
    for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(c + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
#pragma unroll 1
    for (int p = 0; p < n_rep; ++p) {                          // (quantised fp32 output: not a configuration the host issues)
      float* cp = reinterpret_cast<float*>(e.rep[p]) + off;
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(cp + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
  }
}

// Rotary position embedding + matmul_0 operand quantiser on TWO 32-column chunks of one accumulator row that are half a head apart
// (features e0 .. e0+31 and e0 + d/2 .. e0 + d/2 + 31 of one head): replaces the fp32 store of q_proj / k_proj, the ~12 element-wise
// torch kernels of the reference's apply_rotary_pos_emb (models/llama_quantized/modeling_llama.py:309-314 via
// quantized_functions/rotary_positional_encoding.py:27-36) and the x / y quantizers of matmul_0 (quantized_functions/matmul.py:165-193).
// Same arithmetic and order as rope_quant_q/k_kernel (quantize.cu), hence the same bits: rn(rn(x * cos) + rn(rot * sin)),
// rot = -x[i + d/2] in the lower half and x[i - d/2] in the upper one.
// 32 table rows (one per lane: row p_lane of a [rows][d] fp32 table) x 32 columns starting at col -> row-per-lane registers.  Loaded as
// 8 instructions of 4 full 128-byte row segments (the row index of the segment's owner comes by shuffle) and transposed through the warp's
// scratch — the inverse of chunk_to_coalesced.  (Lane-per-row 16-byte loads touched 32 lines per instruction: the epilogue then took
// longer than its mainloop — 0.133 ms against 0.092 ms for the q_proj GEMM of a Llama-7B layer.)
__device__ __forceinline__ void load_table_rows32(const float* table, int p_lane, int d, int col, float (&o)[32], uint32_t* scratch, int lane) {
  float4* sc = reinterpret_cast<float4*>(scratch);
  const int rsub = lane >> 3, c = lane & 7;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = i * 4 + rsub;
    const int pr = __shfl_sync(0xffffffffu, p_lane, r);
    sc[r * 8 + (c ^ (r & 7))] = __ldg(reinterpret_cast<const float4*>(table + (int64_t)pr * d + col) + c);
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 t = sc[lane * 8 + (j ^ (lane & 7))];
    o[4 * j] = t.x; o[4 * j + 1] = t.y; o[4 * j + 2] = t.z; o[4 * j + 3] = t.w;
  }
  __syncwarp();                                                  // scratch is reused by the next strip
}

__device__ __forceinline__ void epilogue_rope_pair(const GemmArgs& g, uint32_t (&lo)[32], uint32_t (&hi)[32], int row, int col_lo, int col_hi,
                                                   bool row_ok, uint32_t* scratch, int lane, float* Cb) {
  const EpiArgs& e = g.epi;
  const int d = e.rope_d;
  // segment of this pair (warp-uniform): its destination, block direction, format and whether it is rotated at all (v_proj is not)
  int qmode = e.qmode, rope = 1;
  const FmtParams* qf = &e.q;
  if (e.seg_cols > 0) {
    const int seg = col_lo / e.seg_cols;
    qmode = e.seg_qmode[seg];
    rope = e.seg_rope[seg];
    qf = &e.seg_q[seg];
    Cb = reinterpret_cast<float*>(e.seg_C[seg]);
  }
  const int gcol_lo = col_lo, gcol_hi = col_hi;                // the bias keeps the global (concatenated) column index
  if (e.seg_cols > 0) {                                        // column inside the segment's own output
    col_lo %= e.seg_cols;
    col_hi %= e.seg_cols;
  }
  int p = 0;                                                   // table row of this lane's token (host: rope_rows < 2^31)
  if (row_ok && rope) p = e.rope_pos ? (int)min(max(e.rope_pos[row], (int64_t)0), e.rope_rows - 1) : (row % e.rope_S);
  const int e0 = col_lo % d;                                   // position of the lower chunk inside its head (< d / 2)
  float a[32], b[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) { a[j] = __uint_as_float(lo[j]); b[j] = __uint_as_float(hi[j]); }
  if (g.bias) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 b0 = *reinterpret_cast<const float4*>(g.bias + gcol_lo + j), b1 = *reinterpret_cast<const float4*>(g.bias + gcol_hi + j);
      a[j] = __fadd_rn(a[j], b0.x); a[j + 1] = __fadd_rn(a[j + 1], b0.y); a[j + 2] = __fadd_rn(a[j + 2], b0.z); a[j + 3] = __fadd_rn(a[j + 3], b0.w);
      b[j] = __fadd_rn(b[j], b1.x); b[j + 1] = __fadd_rn(b[j + 1], b1.y); b[j + 2] = __fadd_rn(b[j + 2], b1.z); b[j + 3] = __fadd_rn(b[j + 3], b1.w);
    }
  }
  if (rope) {
    // products first (x * cos of both halves), then the sine terms: at most one table strip is live at a time
    float t[32], xs[32];
    load_table_rows32(e.rope_cos, p, d, e0, t, scratch, lane);
#pragma unroll
    for (int j = 0; j < 32; ++j) { xs[j] = a[j]; a[j] = __fmul_rn(a[j], t[j]); }
    load_table_rows32(e.rope_sin, p, d, e0, t, scratch, lane);
#pragma unroll
    for (int j = 0; j < 32; ++j) a[j] = __fadd_rn(a[j], __fmul_rn(-b[j], t[j]));
    load_table_rows32(e.rope_cos, p, d, e0 + (d >> 1), t, scratch, lane);
#pragma unroll
    for (int j = 0; j < 32; ++j) b[j] = __fmul_rn(b[j], t[j]);
    load_table_rows32(e.rope_sin, p, d, e0 + (d >> 1), t, scratch, lane);
#pragma unroll
    for (int j = 0; j < 32; ++j) b[j] = __fadd_rn(b[j], __fmul_rn(xs[j], t[j]));
  }
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = half ? b[j] : a[j];
    if (qmode == 1) {
#pragma unroll
      for (int blk = 0; blk < 2; ++blk) {
        float t[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) t[i] = v[blk * 16 + i];
        quantize_signed16_rt(t, *qf);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[blk * 16 + i] = t[i];
      }
    } else {
      // blocks of 16 consecutive rows (k^T operand): every lane takes part; M % 16 == 0 keeps a block all-valid or all-invalid
      if (qf->kind == kBlockFP) quant_rowblocks32<kBlockFP>(v, *qf, scratch, lane);
      else quant_rowblocks32<kBlockMinifloat>(v, *qf, scratch, lane);
    }
    if (row_ok) {
      __nv_bfloat16* c = reinterpret_cast<__nv_bfloat16*>(Cb) + (int64_t)row * g.ldc + (half ? col_hi : col_lo);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        reinterpret_cast<uint4*>(c)[j] = make_uint4(pack_bf16_rn(v[8 * j], v[8 * j + 1]), pack_bf16_rn(v[8 * j + 2], v[8 * j + 3]),
                                                    pack_bf16_rn(v[8 * j + 4], v[8 * j + 5]), pack_bf16_rn(v[8 * j + 6], v[8 * j + 7]));
    }
  }
}

template <int BN, int EPI, int CG>
__global__ void __launch_bounds__(GemmCfg<BN, EPI, CG>::kThreads, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmArgs g) {
  using Cfg = GemmCfg<BN, EPI, CG>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t smem_base = ptx::smem_u32(smem);
  const uint32_t bar_base = smem_base + Cfg::kStages * Cfg::kStage;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::kStages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::kStages + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + Cfg::kStages * Cfg::kStage + 8 * (2 * Cfg::kStages + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = (CG == 2) ? (int)ptx::cluster_ctarank() : 0;      // 0 = leader (issues the MMAs)

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(tfull_bar(a), 1);
      ptx::mbar_init(tempty_bar(a), CG * Cfg::kEpiWarps);      // one arrive per epilogue warp (of both CTAs of a pair)
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    if (CG == 2) ptx::tmem_alloc_cg2<Cfg::kTmemCols>(tmem_slot);
    else ptx::tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  }
  ptx::tc_fence_before();
  if (CG == 2) ptx::cluster_sync(); else __syncthreads();     // the peer's barriers exist before anything signals them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // programmatic dependent launch: everything above ran while the previous kernel of the stream was draining; from here on
  // global memory is read and written, which needs that kernel COMPLETE.  The next kernel may start its own set-up as soon as
  // SMs free up (it blocks in the same wait until this grid has finished).
  ptx::griddep_launch();
  ptx::griddep_wait();

  const int kb_per_term = (g.K + kBK - 1) / kBK;
  const int num_kb = kb_per_term * (g.n_terms > 0 ? g.n_terms : 1);
  const int tiles_per_batch = g.tiles_m * g.tiles_n;          // CG == 2: tiles_m counts 256-row blocks
  const int total_tiles = tiles_per_batch * g.batch;
  const int worker = blockIdx.x / CG, num_workers = gridDim.x / CG;
  // tile -> (batch, first row of THIS CTA, n block)
  auto decode = [&](int tile, int& b, int& row0, int& nb) {
    b = tile / tiles_per_batch;
    const int t = tile - b * tiles_per_batch;
    int mb;
    if (g.band_m > 0) {
      // banded raster for a B operand that cannot stay in L2 (lm_head: 412 MB of planes): a band's A rows (<= 32 MB) stay in L2
      // while the N blocks stream past once per band; the ~74 tiles in flight share 16 A blocks and ~5 B blocks
      const int per_band = g.band_m * g.tiles_n;
      const int band = t / per_band, r = t - band * per_band, m0 = band * g.band_m;
      const int bm = min(g.band_m, g.tiles_m - m0);
      nb = r / bm;
      mb = m0 + (r - nb * bm);
    } else {
      mb = t / g.tiles_n;                                      // n fastest: the workers of a wave share A rows through L2
      nb = t - mb * g.tiles_n;
    }
    row0 = (mb * CG + rank) * kBM;
  };

  // causal modes: first row past this tile (of the CTA pair's 256 rows when CG == 2) decides what is skipped
  auto tile_skipped = [&](int row0, int nb) { return g.causal == 1 && nb * BN >= (row0 - rank * kBM) + CG * kBM; };
  auto tile_kb_per_term = [&](int row0) {
    return g.causal == 2 ? min(kb_per_term, ((row0 - rank * kBM) + CG * kBM + kBK - 1) / kBK) : kb_per_term;
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = worker; tile < total_tiles; tile += num_workers) {
        int b, row0, nb;
        decode(tile, b, row0, nb);
        if (tile_skipped(row0, nb)) continue;
        const int kpt = tile_kb_per_term(row0), nkb = kpt * (g.n_terms > 0 ? g.n_terms : 1);
        for (int kb = 0; kb < nkb; ++kb) {
          ptx::mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * Cfg::kStage;
          int ca = b, cb = g.b_broadcast ? 0 : b, kk = kb;
          if (g.n_terms > 0) {
            const int t = kb / kpt;
            kk = kb - t * kpt;
            ca = g.term_a[t] * g.batch + b;                       // planes are stacked outside the batch: [plane][batch][rows][K]
            cb = g.term_b[t] * (g.b_broadcast ? 1 : g.batch) + (g.b_broadcast ? 0 : b);
          }
          if (CG == 2) {
            // the leader's barrier collects the bytes of BOTH CTAs' loads
            if (rank == 0) ptx::mbar_expect_tx(full_bar(stage), 2 * Cfg::kStage);
            ptx::tma_load_3d_cg2(sa, &tmA, full_bar(stage), kk * kBK, row0, ca);
            ptx::tma_load_3d_cg2(sa + Cfg::kStageA, &tmB, full_bar(stage), kk * kBK, nb * BN + rank * (BN / 2), cb);
          } else {
            ptx::mbar_expect_tx(full_bar(stage), Cfg::kStage);
            ptx::tma_load_3d(sa, &tmA, full_bar(stage), kk * kBK, row0, ca);
            ptx::tma_load_3d(sa + Cfg::kStageA, &tmB, full_bar(stage), kk * kBK, nb * BN, cb);
          }
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = g.fp16 ? ptx::idesc_f16_f32(kBM * CG, BN) : ptx::idesc_bf16_f32(kBM * CG, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = worker; tile < total_tiles; tile += num_workers) {
        int nkb = num_kb;
        if (g.causal) {
          int b, row0, nb;
          decode(tile, b, row0, nb);
          if (tile_skipped(row0, nb)) continue;
          nkb = tile_kb_per_term(row0) * (g.n_terms > 0 ? g.n_terms : 1);
        }
        ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < nkb; ++kb) {
          ptx::mbar_wait(full_bar(stage), phase);
          ptx::tc_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::kStage;
          const uint64_t adesc = ptx::smem_desc_sw128_kmajor(sa);
          const uint64_t bdesc = ptx::smem_desc_sw128_kmajor(sa + Cfg::kStageA);
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k) {
            // advance 16 bf16 = 32 bytes along K inside the swizzle atom: +2 in the (addr >> 4) field
            if (CG == 2) ptx::umma_bf16_cg2(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
            else ptx::umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
          }
          // frees the smem slot (in both CTAs of a pair) when these MMAs retire
          if (CG == 2) ptx::umma_commit_cg2_mc(empty_bar(stage), 3); else ptx::umma_commit(empty_bar(stage));
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        // accumulator complete -> epilogue warps (of both CTAs)
        if (CG == 2) ptx::umma_commit_cg2_mc(tfull_bar(acc), 3); else ptx::umma_commit(tfull_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    const int ew = warp - 4;                       // epilogue warp index
    const int q = ew & 3;                          // TMEM lane quarter this warp may access (== warp % 4)
    const int c_first = ew >> 2;                   // with 8 epilogue warps the two warps of a quarter take alternate column chunks
    constexpr int c_step = Cfg::kEpiWarps / 4;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = worker; tile < total_tiles; tile += num_workers) {
      int b, row0, nb;
      decode(tile, b, row0, nb);
      if (tile_skipped(row0, nb)) continue;
      ptx::mbar_wait(tfull_bar(acc), acc_phase);
      ptx::tc_fence_after();
      const int row = row0 + q * 32 + lane;
      if (EPI == 6) {
        // RoPE + quantiser: the unit of work is a PAIR of chunks half a head apart (host: rope_d % 64 == 0, BN % rope_d == 0, N % rope_d == 0)
        uint32_t* scratch = reinterpret_cast<uint32_t*>(smem + Cfg::kEpiOff) + ew * kEpiWarpWords;
        const int hc = g.epi.rope_d >> 6;             // chunks per half head
#pragma unroll 1
        for (int pi = c_first; pi < BN / 64; pi += c_step) {
          const int c_lo = (pi / hc) * 2 * hc + (pi % hc), c_hi = c_lo + hc;
          const int col_lo = nb * BN + c_lo * 32, col_hi = nb * BN + c_hi * 32;
          if (col_lo >= g.N) continue;                // warp-uniform (chunk order is not monotonic in pi for hc > 1)
          uint32_t rl[32], rh[32];
          ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c_lo * 32), rl);
          ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c_hi * 32), rh);
          ptx::tmem_ld_wait();
          epilogue_rope_pair(g, rl, rh, row, col_lo, col_hi, row < g.M, scratch, lane, g.C);
        }
      } else if (EPI) {
        // fused epilogue (batch == 1, N % 32 == 0, all pointers 16-byte aligned: checked on the host)
        uint32_t* scratch = reinterpret_cast<uint32_t*>(smem + Cfg::kEpiOff) + ew * kEpiWarpWords;
        const bool row_ok = row < g.M;
        // row-per-lane residual prefetch: only the generic instance's quantising paths use it (the coalesced path reads it itself)
        const bool has_res = EPI == 1 && g.epi.residual != nullptr && row_ok && !(g.epi.qmode == 0 && !g.epi.out_bf16);
        const float* rrow = has_res ? g.epi.residual + (int64_t)row * g.epi.ldr + nb * BN : nullptr;
        float4 res_next[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) res_next[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (has_res && nb * BN + c_first * 32 < g.N) {
#pragma unroll
          for (int j = 0; j < 8; ++j) res_next[j] = *reinterpret_cast<const float4*>(rrow + c_first * 32 + 4 * j);
        }
#pragma unroll 1
        for (int c = c_first; c < BN / 32; c += c_step) {
          const int col0 = nb * BN + c * 32;
          if (col0 >= g.N) break;                 // warp-uniform
          uint32_t r[32];
          ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c * 32), r);
          float4 res[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) res[j] = res_next[j];
          if (has_res && c + c_step < BN / 32 && col0 + c_step * 32 < g.N) {   // next chunk's residual: in flight during this chunk's math
#pragma unroll
            for (int j = 0; j < 8; ++j) res_next[j] = *reinterpret_cast<const float4*>(rrow + (c + c_step) * 32 + 4 * j);
          }
          ptx::tmem_ld_wait();
          epilogue_chunk<EPI>(g, r, res, row, col0, row_ok, scratch, lane, g.C + (int64_t)b * g.sc);      // batch b of a batched problem
        }
      } else {
        float* crow = g.C + (int64_t)b * g.sc + (int64_t)row * g.ldc;
        const bool vec_ok = ((g.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0) && ((g.sc & 3) == 0) &&
                            ((reinterpret_cast<uintptr_t>(g.bias) & 15) == 0) && ((reinterpret_cast<uintptr_t>(g.col_scale) & 15) == 0) &&
                            (g.batch == 1 || (g.N & 3) == 0);
#pragma unroll 1
        for (int c = c_first; c < BN / 32; c += c_step) {
          uint32_t r[32];
          ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c * 32), r);
          ptx::tmem_ld_wait();
          const int col0 = nb * BN + c * 32;
          if (row < g.M && col0 < g.N) {
            if (vec_ok && col0 + 32 <= g.N) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                float4 o = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                       __uint_as_float(r[j + 3]));
                if (g.row_scale) {
                  const float rs = g.row_scale[(int64_t)b * g.M + row];
                  const float4 cs = *reinterpret_cast<const float4*>(g.col_scale + (g.b_broadcast ? 0 : (int64_t)b * g.N) + col0 + j);
                  o.x *= rs * cs.x; o.y *= rs * cs.y; o.z *= rs * cs.z; o.w *= rs * cs.w;
                }
                if (g.bias) {
                  const float4 bv = *reinterpret_cast<const float4*>(g.bias + col0 + j);
                  o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
                }
                *reinterpret_cast<float4*>(crow + col0 + j) = o;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                if (col0 + j < g.N) {
                  float o = __uint_as_float(r[j]);
                  if (g.row_scale) o *= g.row_scale[(int64_t)b * g.M + row] * g.col_scale[(g.b_broadcast ? 0 : (int64_t)b * g.N) + col0 + j];
                  crow[col0 + j] = o + (g.bias ? g.bias[col0 + j] : 0.f);
                }
              }
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) ptx::mbar_arrive_leader(tempty_bar(acc)); else ptx::mbar_arrive(tempty_bar(acc));
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  ptx::tc_fence_before();
  if (CG == 2) ptx::cluster_sync(); else __syncthreads();     // nobody leaves while the peer may still signal / read this CTA
  if (warp == 2) {
    ptx::tc_fence_after();
    if (CG == 2) ptx::tmem_dealloc_cg2<Cfg::kTmemCols>(tmem_base);
    else ptx::tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;

static int load_encode() {
  if (g_encode) return BQ_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  BQ_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (qres != cudaDriverEntryPointSuccess || !fn) {
    set_last_cuda_error("cuTensorMapEncodeTiled entry point not available", __FILE__, __LINE__);
    return BQ_ERR_CUDA;
  }
  g_encode = (PFN_encodeTiled)fn;
  return BQ_OK;
}

// bf16 [batch][rows][K] K-major tensor map, box {64, box_rows, 1}, 128B swizzle, zero OOB fill
int make_tmap_bf16_kmajor(CUtensorMap* tm, const void* base, int64_t K, int64_t rows, int64_t batch, int64_t ld,
                          int64_t batch_stride, int box_rows) {
  int rc = load_encode();
  if (rc) return rc;
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)(batch > 0 ? batch : 1)};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)(batch > 1 ? batch_stride : rows * ld) * 2};
  if (strides[1] == 0) strides[1] = (cuuint64_t)rows * ld * 2;
  cuuint32_t box[3] = {(cuuint32_t)kBK, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[96];
    snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    set_last_cuda_error(msg, __FILE__, __LINE__);
    return BQ_ERR_CUDA;
  }
  return BQ_OK;
}

// plain 2-D map over 32-bit elements [rows][inner] (fp32: dtype_code 0, uint32: 1), box {box_inner, box_rows}, optional 128B swizzle
// (box_inner * 4 == 128 then) — the raw operands of gemm_xform_sm100.cu: fp32 activation tiles, packed weight groups
int make_tmap_2d(CUtensorMap* tm, const void* base, int dtype_code, int64_t inner, int64_t rows, int64_t row_stride_bytes, int box_inner,
                 int box_rows, int swizzle128) {
  int rc = load_encode();
  if (rc) return rc;
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)row_stride_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(tm, dtype_code == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void*>(base),
                        dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[96];
    snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled(2d) failed with CUresult %d", (int)r);
    set_last_cuda_error(msg, __FILE__, __LINE__);
    return BQ_ERR_CUDA;
  }
  return BQ_OK;
}

// bf16 [B][S][H][d] (token stride ld_tok elements) as a 4-D map ordered {d, H, S, B} (strides ascending),
// box {64, 1, box_rows, 1}: one head's [box_rows tokens x 64] tile lands as 128-byte rows, 128B-swizzled.
int make_tmap_bf16_4d(CUtensorMap* tm, const void* base, int64_t d, int64_t S, int64_t H, int64_t B, int64_t ld_tok,
                      int box_rows) {
  int rc = load_encode();
  if (rc) return rc;
  cuuint64_t dims[4] = {(cuuint64_t)d, (cuuint64_t)H, (cuuint64_t)S, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)d * 2, (cuuint64_t)ld_tok * 2, (cuuint64_t)S * ld_tok * 2};
  cuuint32_t box[4] = {(cuuint32_t)kBK, 1, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[96];
    snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled(4d) failed with CUresult %d", (int)r);
    set_last_cuda_error(msg, __FILE__, __LINE__);
    return BQ_ERR_CUDA;
  }
  return BQ_OK;
}

template <int BN, int EPI, int CG>
static int launch_gemm_cg(const CUtensorMap& tmA, const CUtensorMap& tmB, GemmArgs g, cudaStream_t st, int kern_id) {
  using Cfg = GemmCfg<BN, EPI, CG>;
  static PerDevice<bool> attr_pd;
  bool& attr_set = attr_pd.get();
  if (!attr_set) {
    BQ_CUDA_CHECK(cudaFuncSetAttribute(gemm_bf16_tn_kernel<BN, EPI, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  g.tiles_m = (g.M + kBM * CG - 1) / (kBM * CG);
  g.tiles_n = (g.N + BN - 1) / BN;
  int64_t total = (int64_t)g.tiles_m * g.tiles_n * g.batch;
  if (total > 0x7fffffffll) return BQ_ERR_UNSUPPORTED;
  g.band_m = 0;
  if (g.batch == 1) {
    int pa = 1, pb = 1;
    for (int t = 0; t < g.n_terms; ++t) { pa = std::max(pa, g.term_a[t] + 1); pb = std::max(pb, g.term_b[t] + 1); }
    const int64_t bytes_b = (int64_t)g.N * g.K * 2 * pb;
    const int64_t bytes_a_block = (int64_t)kBM * CG * g.K * 2 * pa;
    if (bytes_b > (40ll << 20) && g.tiles_m > 1)
      g.band_m = (int)std::min<int64_t>(g.tiles_m, std::max<int64_t>(1, (32ll << 20) / bytes_a_block));
  }
  const int grid = CG * (int)std::min<int64_t>(total, num_sms() / CG);
  {
    LaunchScope ls(kern_id, st);
    BQ_CUDA_CHECK(launch_ex(gemm_bf16_tn_kernel<BN, EPI, CG>, grid, Cfg::kThreads, Cfg::kSmemBytes, st, CG, tmA, tmB, g));
  }
  BQ_CUDA_CHECK(cudaGetLastError());
  return BQ_OK;
}

// CTA pairs need BN == 256, one batch and enough rows; `pair` is decided by the callers (use_pairs) BEFORE the B tensor map
// is built, because its box holds BN / CG rows.
static bool g_pairs_enabled = true;
static bool use_pairs(int64_t batch, int64_t M, int64_t N) { return g_pairs_enabled && batch == 1 && N > 128 && M > 128; }
// Tile shape for a problem: {BN, pair}.  Large problems take the 256 x 256 CTA-pair tile (best operand reuse).  A problem whose
// pair tiles do not even fill the 74 pairs once — the per-rank GEMMs of the tensor-parallel layer at 8 GPUs (M 4096, N 512: 32 pair
// tiles, each a 28 us mainloop at K 4096 with 57 % of the chip idle) — is latency-bound by ONE tile's duration: 128 x 128 tiles on
// single CTAs put every SM to work for half as long.  Cost model in units of per-SM tile work, waves x tile area x a measured
// inefficiency factor of the narrower tiles (less operand reuse per shared-memory byte).
static bool g_small_tiles = true;
static void choose_tile(int64_t batch, int64_t M, int64_t N, int* BN, bool* pair) {
  *BN = (N <= 64) ? 64 : (N <= 128 ? 128 : 256);
  *pair = use_pairs(batch, M, N);
  if (!g_small_tiles || N <= 128 || M <= 128) return;
  const int64_t sms = num_sms();
  auto waves = [](int64_t tiles, int64_t workers) { return (tiles + workers - 1) / workers; };
  const int64_t t_pair = ((M + 255) / 256) * ((N + 255) / 256) * batch, t_256 = ((M + 127) / 128) * ((N + 255) / 256) * batch,
                t_128 = ((M + 127) / 128) * ((N + 127) / 128) * batch;
  const double c_pair = *pair ? (double)waves(t_pair, sms / 2) * 32768.0 : 1e30;
  const double c_256 = (double)waves(t_256, sms) * 32768.0 * 1.05, c_128 = (double)waves(t_128, sms) * 16384.0 * 1.25;
  if (c_128 < c_pair && c_128 < c_256) { *BN = 128; *pair = false; }
  else if (c_256 < c_pair) { *BN = 256; *pair = false; }
}

template <bool EPI>
static int launch_gemm_any(int BN, bool pair, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& g, cudaStream_t st,
                           int kern_id) {
  if (EPI && g.epi.rope_cos) {                                   // RoPE + quantiser (BN is a multiple of the head size: checked by the caller)
    if (pair) return launch_gemm_cg<256, 6, 2>(tmA, tmB, g, st, kern_id);
    if (BN == 128) return launch_gemm_cg<128, 6, 1>(tmA, tmB, g, st, kern_id);
    if (BN == 256) return launch_gemm_cg<256, 6, 1>(tmA, tmB, g, st, kern_id);
    return BQ_ERR_UNSUPPORTED;
  }
  if (EPI && g.epi.act == 2) {                                   // gated SiLU: its own instances (N = 2 * features >= 128)
    if (pair) return launch_gemm_cg<256, 5, 2>(tmA, tmB, g, st, kern_id);
    if (BN == 128) return launch_gemm_cg<128, 5, 1>(tmA, tmB, g, st, kern_id);
    if (BN == 256) return launch_gemm_cg<256, 5, 1>(tmA, tmB, g, st, kern_id);
    return BQ_ERR_UNSUPPORTED;
  }
  if (pair) {
    if (!EPI) return launch_gemm_cg<256, 0, 2>(tmA, tmB, g, st, kern_id);
    const EpiArgs& e = g.epi;
    if (e.qmode == 0 && !e.out_bf16) return launch_gemm_cg<256, 2, 2>(tmA, tmB, g, st, kern_id);
    if (e.out_bf16 && !e.residual && e.n_rep == 0) {
      if (e.qmode == 1) return launch_gemm_cg<256, 3, 2>(tmA, tmB, g, st, kern_id);
      if (e.qmode == 2) return launch_gemm_cg<256, 4, 2>(tmA, tmB, g, st, kern_id);
    }
    return launch_gemm_cg<256, 1, 2>(tmA, tmB, g, st, kern_id);
  }
  switch (BN) {
    case 64: return launch_gemm_cg<64, EPI ? 1 : 0, 1>(tmA, tmB, g, st, kern_id);
    case 128: return launch_gemm_cg<128, EPI ? 1 : 0, 1>(tmA, tmB, g, st, kern_id);
    default: return launch_gemm_cg<256, EPI ? 1 : 0, 1>(tmA, tmB, g, st, kern_id);
  }
}

int gemm_bf16_tn_impl(const void* A, const void* B, float* C, const float* bias, int64_t batch, int64_t M, int64_t N,
                      int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int64_t sa, int64_t sb, int64_t sc,
                      cudaStream_t st) {
  if (batch < 0 || M < 0 || N < 0 || K < 0) return BQ_ERR_BAD_ARG;
  if (batch == 0 || M == 0 || N == 0) return BQ_OK;
  if (!A || !B || !C) return BQ_ERR_BAD_ARG;
  if (K == 0) return BQ_ERR_UNSUPPORTED;
  if ((lda % 8) || (ldb % 8) || ((uintptr_t)A % 16) || ((uintptr_t)B % 16) || ((uintptr_t)C % 4)) return BQ_ERR_BAD_ARG;
  if (batch > 1 && ((sa % 8) || (sb % 8))) return BQ_ERR_BAD_ARG;
  if (lda < K || ldb < K || ldc < N) return BQ_ERR_BAD_ARG;
  if (M > 0x7fffffff || N > 0x7fffffff || K > 0x7fffffff || batch > 0x7fffffff) return BQ_ERR_UNSUPPORTED;
  int BN;
  bool pair;
  choose_tile(batch, M, N, &BN, &pair);
  CUtensorMap tmA, tmB;
  int rc = make_tmap_bf16_kmajor(&tmA, A, K, M, batch, lda, sa, kBM);
  if (rc) return rc;
  const bool bcast = (sb == 0) || batch == 1;
  rc = make_tmap_bf16_kmajor(&tmB, B, K, N, bcast ? 1 : batch, ldb, sb, pair ? 128 : BN);
  if (rc) return rc;
  GemmArgs g;
  g.C = C; g.bias = bias; g.M = (int)M; g.N = (int)N; g.K = (int)K; g.batch = (int)batch;
  g.ldc = ldc; g.sc = sc; g.tiles_m = g.tiles_n = 0; g.b_broadcast = bcast ? 1 : 0;
  g.n_terms = 0;
  g.fp16 = 0; g.row_scale = g.col_scale = nullptr; g.causal = 0;
  memset(&g.epi, 0, sizeof(g.epi));
  return launch_gemm_any<false>(BN, pair, tmA, tmB, g, st, kKernGemm);
}

int make_params(const bq_format* f, FmtParams* p);

// GEMM with the fused epilogue (see EpiArgs).  C: fp32 or bf16 [M][ldc].
int gemm_bf16_tn_epi_impl(const void* A, const void* B, void* C, const bq_gemm_epilogue* ep, int64_t M, int64_t N, int64_t K,
                          int64_t lda, int64_t ldb, int64_t ldc, cudaStream_t st) {
  if (!ep || M < 0 || N < 0 || K < 0) return BQ_ERR_BAD_ARG;
  if (M == 0 || N == 0) return BQ_OK;
  if (!A || !B || !C) return BQ_ERR_BAD_ARG;
  if (K == 0) return BQ_ERR_UNSUPPORTED;
  if ((lda % 8) || (ldb % 8) || ((uintptr_t)A % 16) || ((uintptr_t)B % 16) || ((uintptr_t)C % 16)) return BQ_ERR_BAD_ARG;
  if (lda < K || ldb < K || ldc < (ep->act == 2 ? N / 2 : N)) return BQ_ERR_BAD_ARG;
  if (M > 0x7fffffff || N > 0x7fffffff || K > 0x7fffffff) return BQ_ERR_UNSUPPORTED;
  if (N % 32) return BQ_ERR_UNSUPPORTED;
  const bool out_bf16 = ep->out_dtype == BQ_BF16;
  if (ep->out_dtype != BQ_F32 && !out_bf16) return BQ_ERR_BAD_ARG;
  if (ldc % (out_bf16 ? 8 : 4)) return BQ_ERR_BAD_ARG;
  if (ep->bias && ((uintptr_t)ep->bias % 16)) return BQ_ERR_BAD_ARG;
  if (ep->residual && (((uintptr_t)ep->residual % 16) || (ep->ldr % 4) || ep->ldr < N)) return BQ_ERR_BAD_ARG;
  if (ep->act != 0 && ep->act != 1 && ep->act != 2) return BQ_ERR_UNSUPPORTED;
  if (ep->act == 2) {
    // gated SiLU: N counts the interleaved gate / up columns, C is bf16 [M][N / 2] in the format of the consuming Linear
    if (!ep->qfmt || ep->qdir != 0 || !out_bf16 || ep->residual || ep->n_replicas != 0 || ep->scale != 1.0f) return BQ_ERR_UNSUPPORTED;
    if (N < 128 || ldc < N / 2) return BQ_ERR_BAD_ARG;
  }
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.epi.residual = ep->residual; g.epi.ldr = ep->ldr; g.epi.scale = ep->scale; g.epi.act = ep->act;
  g.epi.out_bf16 = out_bf16 ? 1 : 0;
  if (ep->qfmt) {
    const bq_format* f = ep->qfmt;
    if (f->kind != BQ_KIND_BLOCK_FP && f->kind != BQ_KIND_BLOCK_MINIFLOAT && f->kind != BQ_KIND_BLOCK_LOG) return BQ_ERR_UNSUPPORTED;
    if (f->block_rows != 1 || f->block_cols != 16) return BQ_ERR_UNSUPPORTED;     // block of 16 along the chosen direction
    if (ep->qdir != 0 && ep->qdir != 1) return BQ_ERR_BAD_ARG;
    if (f->kind == BQ_KIND_BLOCK_LOG && ep->qdir != 0) return BQ_ERR_UNSUPPORTED;   // (block_log matmuls leave the k^T operand unquantised)
    if (ep->qdir == 1 && (M % 16)) return BQ_ERR_UNSUPPORTED;
    int rc = make_params(f, &g.epi.q);
    if (rc) return rc;
    g.epi.q.fold_zero = 0;
    g.epi.qmode = ep->qdir == 1 ? 2 : 1;
  }
  if (ep->n_replicas < 0 || ep->n_replicas > BQ_MAX_REPLICAS) return BQ_ERR_BAD_ARG;
  g.epi.n_rep = ep->n_replicas;
  for (int i = 0; i < ep->n_replicas; ++i) {
    if (!ep->replicas[i] || ((uintptr_t)ep->replicas[i] % 16)) return BQ_ERR_BAD_ARG;
    g.epi.rep[i] = ep->replicas[i];
  }
  int BN;
  bool pair;
  choose_tile(1, M, N, &BN, &pair);
  CUtensorMap tmA, tmB;
  int rc = make_tmap_bf16_kmajor(&tmA, A, K, M, 1, lda, 0, kBM);
  if (rc) return rc;
  rc = make_tmap_bf16_kmajor(&tmB, B, K, N, 1, ldb, 0, pair ? 128 : BN);
  if (rc) return rc;
  g.C = (float*)C; g.bias = ep->bias; g.M = (int)M; g.N = (int)N; g.K = (int)K; g.batch = 1;
  g.ldc = ldc; g.sc = 0; g.b_broadcast = 1; g.n_terms = 0;
  return launch_gemm_any<true>(BN, pair, tmA, tmB, g, st, kKernGemmEpi);
}

// q_proj / k_proj of a Llama layer with the rotary embedding and matmul_0's operand quantiser in the epilogue (see epilogue_rope_pair).
int gemm_bf16_tn_rope_impl(const void* A, const void* B, void* C, const float* bias, const bq_format* qfmt, int qdir, const float* cos_t,
                           const float* sin_t, const int64_t* pos, int64_t table_rows, int64_t S, int64_t d, int64_t M, int64_t N, int64_t K,
                           int64_t lda, int64_t ldb, int64_t ldc, cudaStream_t st) {
  if (!qfmt || M < 0 || N < 0 || K < 0 || S <= 0 || d <= 0) return BQ_ERR_BAD_ARG;
  if (M == 0 || N == 0) return BQ_OK;
  if (!A || !B || !C || !cos_t || !sin_t) return BQ_ERR_BAD_ARG;
  if (K == 0) return BQ_ERR_UNSUPPORTED;
  if ((lda % 8) || (ldb % 8) || (ldc % 8) || ((uintptr_t)A % 16) || ((uintptr_t)B % 16) || ((uintptr_t)C % 16) || ((uintptr_t)cos_t % 16) ||
      ((uintptr_t)sin_t % 16) || (bias && ((uintptr_t)bias % 16)))
    return BQ_ERR_BAD_ARG;
  if (lda < K || ldb < K || ldc < N) return BQ_ERR_BAD_ARG;
  if (M > 0x7fffffff || N > 0x7fffffff || K > 0x7fffffff) return BQ_ERR_UNSUPPORTED;
  if (table_rows < 1 || table_rows > 0x7fffffff || (!pos && table_rows < S)) return BQ_ERR_BAD_ARG;
  if ((d != 64 && d != 128) || (N % d)) return BQ_ERR_UNSUPPORTED;      // a 128- or 256-column tile holds whole heads
  if (qfmt->kind != BQ_KIND_BLOCK_FP && qfmt->kind != BQ_KIND_BLOCK_MINIFLOAT) return BQ_ERR_UNSUPPORTED;
  if (qfmt->block_rows != 1 || qfmt->block_cols != 16) return BQ_ERR_UNSUPPORTED;
  if (qdir != 0 && qdir != 1) return BQ_ERR_BAD_ARG;
  if (qdir == 1 && ((M % 16) || (S % 16))) return BQ_ERR_UNSUPPORTED;   // blocks of 16 tokens never straddle a sequence
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  int rc = make_params(qfmt, &g.epi.q);
  if (rc) return rc;
  g.epi.q.fold_zero = 0;
  g.epi.qmode = qdir == 1 ? 2 : 1;
  g.epi.out_bf16 = 1;
  g.epi.scale = 1.0f;
  g.epi.rope_cos = cos_t; g.epi.rope_sin = sin_t; g.epi.rope_pos = pos; g.epi.rope_rows = table_rows;
  g.epi.rope_S = (int)S; g.epi.rope_d = (int)d;
  int BN;
  bool pair;
  choose_tile(1, M, N, &BN, &pair);
  if (BN < 128) { BN = 128; pair = false; }                              // N is a multiple of d >= 64; the 64-wide instance has no RoPE form
  CUtensorMap tmA, tmB;
  rc = make_tmap_bf16_kmajor(&tmA, A, K, M, 1, lda, 0, kBM);
  if (rc) return rc;
  rc = make_tmap_bf16_kmajor(&tmB, B, K, N, 1, ldb, 0, pair ? 128 : BN);
  if (rc) return rc;
  g.C = (float*)C; g.bias = bias; g.M = (int)M; g.N = (int)N; g.K = (int)K; g.batch = 1;
  g.ldc = ldc; g.sc = 0; g.b_broadcast = 1; g.n_terms = 0;
  return launch_gemm_any<true>(BN, pair, tmA, tmB, g, st, kKernGemmEpi);
}

// q_proj | k_proj | v_proj of a Llama layer as ONE GEMM over the concatenated weights [3 * Hs][K] (same x operand): segment 0 = q (RoPE,
// blocks along the features), 1 = k (RoPE, blocks of 16 tokens), 2 = v (no RoPE, blocks along the features) — three bf16 outputs.
// 768 tiles in one persistent launch fill the last wave better than three launches of 256 (3 x 4 waves -> 11) and expose one
// epilogue tail instead of three.
int gemm_bf16_tn_qkv_rope_impl(const void* A, const void* B, void* Cq, void* Ck, void* Cv, const float* bias, const bq_format* fq,
                               const bq_format* fk, const bq_format* fv, const float* cos_t, const float* sin_t, const int64_t* pos,
                               int64_t table_rows, int64_t S, int64_t d, int64_t M, int64_t Hs, int64_t K, int64_t lda, int64_t ldb,
                               int64_t ldc, cudaStream_t st) {
  if (!fq || !fk || !fv || M < 0 || Hs < 0 || K < 0 || S <= 0 || d <= 0) return BQ_ERR_BAD_ARG;
  if (M == 0 || Hs == 0) return BQ_OK;
  if (!A || !B || !Cq || !Ck || !Cv || !cos_t || !sin_t) return BQ_ERR_BAD_ARG;
  if (K == 0) return BQ_ERR_UNSUPPORTED;
  if ((lda % 8) || (ldb % 8) || (ldc % 8) || ((uintptr_t)A % 16) || ((uintptr_t)B % 16) || ((uintptr_t)Cq % 16) || ((uintptr_t)Ck % 16) ||
      ((uintptr_t)Cv % 16) || ((uintptr_t)cos_t % 16) || ((uintptr_t)sin_t % 16) || (bias && ((uintptr_t)bias % 16)))
    return BQ_ERR_BAD_ARG;
  if (lda < K || ldb < K || ldc < Hs) return BQ_ERR_BAD_ARG;
  const int64_t N = 3 * Hs;
  if (M > 0x7fffffff || N > 0x7fffffff || K > 0x7fffffff) return BQ_ERR_UNSUPPORTED;
  if (table_rows < 1 || table_rows > 0x7fffffff || (!pos && table_rows < S)) return BQ_ERR_BAD_ARG;
  if ((d != 64 && d != 128) || (Hs % d) || (Hs % 256)) return BQ_ERR_UNSUPPORTED;      // a tile never straddles two segments
  if ((M % 16) || (S % 16)) return BQ_ERR_UNSUPPORTED;
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  const bq_format* fs[3] = {fq, fk, fv};
  for (int i = 0; i < 3; ++i) {
    if (fs[i]->kind != BQ_KIND_BLOCK_FP && fs[i]->kind != BQ_KIND_BLOCK_MINIFLOAT) return BQ_ERR_UNSUPPORTED;
    if (fs[i]->block_rows != 1 || fs[i]->block_cols != 16) return BQ_ERR_UNSUPPORTED;
    int rc = make_params(fs[i], &g.epi.seg_q[i]);
    if (rc) return rc;
    g.epi.seg_q[i].fold_zero = 0;
  }
  g.epi.seg_cols = (int)Hs;
  g.epi.seg_C[0] = Cq; g.epi.seg_C[1] = Ck; g.epi.seg_C[2] = Cv;
  g.epi.seg_qmode[0] = 1; g.epi.seg_qmode[1] = 2; g.epi.seg_qmode[2] = 1;
  g.epi.seg_rope[0] = 1; g.epi.seg_rope[1] = 1; g.epi.seg_rope[2] = 0;
  g.epi.q = g.epi.seg_q[0];
  g.epi.qmode = 1;
  g.epi.out_bf16 = 1;
  g.epi.scale = 1.0f;
  g.epi.rope_cos = cos_t; g.epi.rope_sin = sin_t; g.epi.rope_pos = pos; g.epi.rope_rows = table_rows;
  g.epi.rope_S = (int)S; g.epi.rope_d = (int)d;
  int BN;
  bool pair;
  choose_tile(1, M, N, &BN, &pair);
  if (BN < 128) { BN = 128; pair = false; }
  CUtensorMap tmA, tmB;
  int rc = make_tmap_bf16_kmajor(&tmA, A, K, M, 1, lda, 0, kBM);
  if (rc) return rc;
  rc = make_tmap_bf16_kmajor(&tmB, B, K, N, 1, ldb, 0, pair ? 128 : BN);
  if (rc) return rc;
  g.C = (float*)Cq; g.bias = bias; g.M = (int)M; g.N = (int)N; g.K = (int)K; g.batch = 1;
  g.ldc = ldc; g.sc = 0; g.b_broadcast = 1; g.n_terms = 0;
  return launch_gemm_any<true>(BN, pair, tmA, tmB, g, st, kKernGemmEpi);
}

// Split-precision GEMM: C = sum over terms (A_plane[ta] @ B_plane[tb]^T) (+ bias).  With x = x0 + x1 + x2 (three bf16
// planes, see split3 in quantize.cu) and the six terms {(2,0),(0,2),(1,1),(1,0),(0,1),(0,0)} (smallest first) the
// result carries ~2^-24 relative error per product, i.e. it stands in for an fp32 GEMM on the tensor cores.
int gemm_split_tn_impl(const void* A, const void* B, float* C, const float* bias, int64_t M, int64_t N, int64_t K,
                       int planes_a, int planes_b, int n_terms, const int* ta, const int* tb, int64_t ldc, cudaStream_t st,
                       int fp16 = 0, const float* row_scale = nullptr, const float* col_scale = nullptr, int64_t batch = 1,
                       int64_t sc = 0, int causal = 0) {
  if (M < 0 || N < 0 || K <= 0 || n_terms < 1 || n_terms > 8 || !ta || !tb || batch < 0) return BQ_ERR_BAD_ARG;
  if (M == 0 || N == 0 || batch == 0) return BQ_OK;
  // C of batch z starts at z * sc: batch-major (sc >= M * ldc) or interleaved inside the rows (ldc >= batch * sc, sc >= N — e.g. the
  // heads of one sequence written straight into the token-major [S][heads * d] activation)
  if (batch > 1 && (!(sc >= M * ldc || (sc >= N && ldc >= batch * sc)) || planes_a * batch > 0x7fffffff || planes_b * batch > 0x7fffffff))
    return BQ_ERR_BAD_ARG;
  if (!A || !B || !C) return BQ_ERR_BAD_ARG;
  if ((K % 8) || ((uintptr_t)A % 16) || ((uintptr_t)B % 16) || ((uintptr_t)C % 4) || ldc < N) return BQ_ERR_BAD_ARG;
  if ((row_scale == nullptr) != (col_scale == nullptr)) return BQ_ERR_BAD_ARG;
  if (M > 0x7fffffff || N > 0x7fffffff || K > 0x7fffffff) return BQ_ERR_UNSUPPORTED;
  for (int i = 0; i < n_terms; ++i)
    if (ta[i] < 0 || ta[i] >= planes_a || tb[i] < 0 || tb[i] >= planes_b) return BQ_ERR_BAD_ARG;
  if (causal < 0 || causal > 2 || (causal == 1 && M != N) || (causal == 2 && M != K)) return BQ_ERR_BAD_ARG;
  int BN;
  bool pair;
  choose_tile(batch, M, N, &BN, &pair);
  CUtensorMap tmA, tmB;
  int rc = make_tmap_bf16_kmajor(&tmA, A, K, M, planes_a * batch, K, M * K, kBM);     // 16-bit elements: the map only moves bytes
  if (rc) return rc;
  rc = make_tmap_bf16_kmajor(&tmB, B, K, N, planes_b * batch, K, N * K, pair ? 128 : BN);
  if (rc) return rc;
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.C = C; g.bias = bias; g.M = (int)M; g.N = (int)N; g.K = (int)K; g.batch = (int)batch;
  g.ldc = ldc; g.sc = batch > 1 ? sc : 0; g.tiles_m = g.tiles_n = 0; g.b_broadcast = batch > 1 ? 0 : 1;
  g.n_terms = n_terms;
  g.fp16 = fp16; g.row_scale = row_scale; g.col_scale = col_scale;
  g.causal = causal;
  for (int i = 0; i < n_terms; ++i) {
    if (ta[i] < 0 || ta[i] >= planes_a || tb[i] < 0 || tb[i] >= planes_b) return BQ_ERR_BAD_ARG;
    g.term_a[i] = (int8_t)ta[i];
    g.term_b[i] = (int8_t)tb[i];
  }
  // fp32 results of the bf16-plane products leave through the coalesced epilogue (32 x 32 chunk transposed in shared memory: full
  // 128-byte row segments per store instruction) when the layout allows; the lane-per-row 16-byte stores of the plain epilogue
  // bound the S x S score writes of the split attention (1 GB per Llama-7B layer)
  if (!fp16 && !row_scale && (N % 32) == 0 && (ldc % 4) == 0 && ((uintptr_t)C % 16) == 0 && (batch == 1 || (sc % 4) == 0) &&
      (!bias || ((uintptr_t)bias % 16) == 0)) {
    g.epi.scale = 1.0f;
    return launch_gemm_any<true>(BN, pair, tmA, tmB, g, st, kKernGemmSplit);
  }
  return launch_gemm_any<false>(BN, pair, tmA, tmB, g, st, kKernGemmSplit);
}

}  // namespace bq

// debugging / measurement switch: 0 forces cta_group::1 tiles everywhere
extern "C" void bq_set_cta_pairs(int on) { bq::g_pairs_enabled = on != 0; }
// A/B switch: 0 = always the largest tile the shape admits (round-1 behaviour), 1 (default) = cost model of choose_tile
extern "C" void bq_set_small_tiles(int on) { bq::g_small_tiles = on != 0; }

extern "C" int bq_gemm_split16_tn(const void* A_planes_f16, const void* B_planes_f16, float* C, const float* bias,
                                  const float* a_inv_scale, const float* b_inv_scale, int64_t M, int64_t N, int64_t K,
                                  int32_t n_terms, const int32_t* term_a, const int32_t* term_b, int64_t ldc, void* stream) {
  if (!a_inv_scale || !b_inv_scale) return BQ_ERR_BAD_ARG;
  return bq::gemm_split_tn_impl(A_planes_f16, B_planes_f16, C, bias, M, N, K, 2, 2, n_terms, term_a, term_b, ldc,
                                (cudaStream_t)stream, 1, a_inv_scale, b_inv_scale);
}

extern "C" int bq_bmm_split16_tn(const void* A_planes_f16, const void* B_planes_f16, float* C, const float* a_inv_scale,
                                const float* b_inv_scale, int64_t batch, int64_t M, int64_t N, int64_t K, int32_t n_terms,
                                const int32_t* term_a, const int32_t* term_b, int64_t ldc, int64_t sc, void* stream) {
  if (!a_inv_scale || !b_inv_scale) return BQ_ERR_BAD_ARG;
  return bq::gemm_split_tn_impl(A_planes_f16, B_planes_f16, C, nullptr, M, N, K, 2, 2, n_terms, term_a, term_b, ldc,
                                (cudaStream_t)stream, 1, a_inv_scale, b_inv_scale, batch, sc);
}

extern "C" int bq_bmm_split_tn(const void* A_planes, const void* B_planes, float* C, int64_t batch, int64_t M, int64_t N, int64_t K,
                               int32_t planes_a, int32_t planes_b, int32_t n_terms, const int32_t* term_a, const int32_t* term_b,
                               int64_t ldc, int64_t sc, int32_t causal, void* stream) {
  return bq::gemm_split_tn_impl(A_planes, B_planes, C, nullptr, M, N, K, planes_a, planes_b, n_terms, term_a, term_b, ldc,
                                (cudaStream_t)stream, 0, nullptr, nullptr, batch, sc, causal);
}

extern "C" int bq_gemm_bf16_tn_ex(const void* A, const void* B, void* C, const bq_gemm_epilogue* ep, int64_t M, int64_t N,
                                  int64_t K, int64_t lda, int64_t ldb, int64_t ldc, void* stream) {
  return bq::gemm_bf16_tn_epi_impl(A, B, C, ep, M, N, K, lda, ldb, ldc, (cudaStream_t)stream);
}

extern "C" int bq_gemm_bf16_tn_rope(const void* A, const void* B, void* C_bf16, const float* bias, const bq_format* qfmt, int32_t qdir,
                                    const float* cos_table, const float* sin_table, const int64_t* position_ids, int64_t table_rows,
                                    int64_t S, int64_t head_dim, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc,
                                    void* stream) {
  return bq::gemm_bf16_tn_rope_impl(A, B, C_bf16, bias, qfmt, qdir, cos_table, sin_table, position_ids, table_rows, S, head_dim, M, N, K,
                                    lda, ldb, ldc, (cudaStream_t)stream);
}

extern "C" int bq_gemm_bf16_tn_qkv_rope(const void* A, const void* B_qkv, void* Cq_bf16, void* Ck_bf16, void* Cv_bf16, const float* bias,
                                        const bq_format* fq, const bq_format* fk, const bq_format* fv, const float* cos_table,
                                        const float* sin_table, const int64_t* position_ids, int64_t table_rows, int64_t S,
                                        int64_t head_dim, int64_t M, int64_t H, int64_t K, int64_t lda, int64_t ldb, int64_t ldc,
                                        void* stream) {
  return bq::gemm_bf16_tn_qkv_rope_impl(A, B_qkv, Cq_bf16, Ck_bf16, Cv_bf16, bias, fq, fk, fv, cos_table, sin_table, position_ids, table_rows,
                                        S, head_dim, M, H, K, lda, ldb, ldc, (cudaStream_t)stream);
}

extern "C" int bq_gemm_split_tn(const void* A_planes, const void* B_planes, float* C, const float* bias, int64_t M, int64_t N,
                                int64_t K, int32_t planes_a, int32_t planes_b, int32_t n_terms, const int32_t* term_a,
                                const int32_t* term_b, int64_t ldc, void* stream) {
  return bq::gemm_split_tn_impl(A_planes, B_planes, C, bias, M, N, K, planes_a, planes_b, n_terms, term_a, term_b, ldc,
                                (cudaStream_t)stream);
}

extern "C" int bq_gemm_bf16_tn(const void* A, const void* B, float* C, const float* bias, int64_t batch, int64_t M,
                               int64_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int64_t sa, int64_t sb,
                               int64_t sc, void* stream) {
  return bq::gemm_bf16_tn_impl(A, B, C, bias, batch, M, N, K, lda, ldb, ldc, sa, sb, sc, (cudaStream_t)stream);
}
