/*
 * bq.h — C ABI of libbq_b200.so: B200 (sm_100a) block-quantisation hot path.
 *
 * This is the drop-in boundary for llm-mixed-q's software-emulated quantisation
 * path.  The reference (ChengZhang-98/llm-mixed-q @ 740bf48) has no FFI: its
 * boundary is the Python operator API under src/llm_mixed_q/models/quantize/.
 * Every entry point below names the reference function it replaces (file:line,
 * relative to that directory).  The Python host mirror (llm_mixed_q_b200/) binds
 * these with ctypes and keeps the reference's names / kwargs / exceptions.
 *
 * Conventions
 *   - plain C types only; every pointer is a DEVICE pointer unless it says "host";
 *   - the caller (PyTorch) owns every buffer, including workspaces — nothing here
 *     allocates or frees device memory;
 *   - every call is asynchronous on the `stream` argument (a cudaStream_t passed as
 *     void*) and re-entrant; the only mutable process-wide state is (a) immutable-after-
 *     first-use caches (TMA descriptors, function attributes), (b) the launch / profile
 *     counters and (c) the explicitly listed A/B switches (bq_set_* near the end of this
 *     file: numerator mode of the attention kernel, kernel-variant overrides).  A switch
 *     is read once per call on the calling thread; change it only between calls;
 *   - return value: 0 on success, otherwise a bq_status (see bq_strerror);
 *   - there is NO CPU fallback: on a machine without a B200 the calls fail with
 *     BQ_ERR_CUDA.
 */
#ifndef BQ_B200_H
#define BQ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BQ_ABI_VERSION 5

#if defined(__GNUC__)
#define BQ_API __attribute__((visibility("default")))
#else
#define BQ_API
#endif

typedef enum bq_status {
  BQ_OK = 0,
  BQ_ERR_BAD_ARG = 1,      /* null pointer, negative size, misaligned pointer          -> ValueError          */
  BQ_ERR_UNSUPPORTED = 2,  /* legal in the reference but not implemented here           -> NotImplementedError */
  BQ_ERR_BAD_FORMAT = 3,   /* width / exponent width outside the representable range    -> ValueError          */
  BQ_ERR_WORKSPACE = 4,    /* workspace too small (see the *_workspace_bytes functions) -> ValueError          */
  BQ_ERR_CUDA = 5,         /* a CUDA runtime/driver call failed                         -> RuntimeError        */
  BQ_ERR_NOT_BF16_EXACT = 6/* quantised operand does not fit bf16's 8 significant bits  -> NotImplementedError */
} bq_status;

/* Arithmetic ("name" key of the reference's quant config, quant_config_parser.py:32-155). */
typedef enum bq_kind {
  BQ_KIND_BLOCK_FP = 0,         /* quantizers/block_fp.py:21-96        */
  BQ_KIND_BLOCK_MINIFLOAT = 1,  /* quantizers/block_minifloat.py:22-74 */
  BQ_KIND_BLOCK_LOG = 2,        /* quantizers/block_log.py:23-69       */
  BQ_KIND_MINIFLOAT_DENORM = 3, /* quantizers/minifloat.py:21-82       */
  BQ_KIND_MINIFLOAT_IEEE = 4,   /* quantizers/minifloat.py:134-196 (scalar bias) */
  BQ_KIND_INTEGER = 5,          /* quantizers/integer.py:25-58         */
  BQ_KIND_NONE = 6              /* bypass: plain fp32 -> bf16 round-to-nearest (GEMM feeds only) */
} bq_kind;

/*
 * One operand format = the `<prefix>_*` keys of a reference quant-config node
 * (prefix = data_in | weight | bias).  The host resolves the reference's
 * `exponent_bias in (None,"none","None") -> 2^(ew-1)-1` rule (block_fp.py:61-62)
 * before filling this struct.
 */
typedef struct bq_format {
  int32_t kind;                 /* bq_kind                                                        */
  int32_t width;                /* *_width                                                        */
  int32_t exponent_width;       /* *_exponent_width   (block_fp, block_minifloat, minifloat_*)     */
  int32_t exponent_bias;        /* *_exponent_bias, resolved (block_fp, minifloat_*); integer: *_frac_width */
  int32_t exponent_bias_width;  /* *_exponent_bias_width (block_minifloat, block_log)             */
  int32_t block_rows;           /* inferred block extent on the second-to-last dim (utils.py:42-67) */
  int32_t block_cols;           /* inferred block extent on the last dim                          */
  int32_t fold_zero;            /* 1: the reference un-blocks this case with F.fold (2-D weight / 3-D
                                   activation, utils.py:186-258), which turns -0.0 into +0.0      */
} bq_format;

/* Output element type of a quantizer launch. */
typedef enum bq_dtype { BQ_F32 = 0, BQ_BF16 = 1 } bq_dtype;

/*
 * Logical operand of a quantizer call: a 3-D fp32 tensor [L, R, C] (the reference's
 * 1-D bias / 2-D activation / 2-D weight / 3-D activation cases canonicalised, L is
 * never blocked), addressed with element strides so that transposed views such as
 * k^T in bmm_0 (models/opt_quantized/modeling_opt.py:246) need no copy.
 */
typedef struct bq_tensor3 {
  int64_t L, R, C;
  int64_t sL, sR, sC;           /* element strides of the INPUT                                    */
} bq_tensor3;

BQ_API const char* bq_strerror(int status);
BQ_API int bq_abi_version(void);
/* last CUDA error string seen by this thread inside the library ("" if none) */
BQ_API const char* bq_last_cuda_error(void);

/* ------------------------------------------------------------------------------------------------
 * Quantizers.  Replaces block_fp_quantizer (block_fp.py:127-153), block_minifloat_quantizer
 * (block_minifloat.py:110-141), block_log_quantizer (block_log.py:95-120),
 * minifloat_denorm_quantizer (minifloat.py:104-131), minifloat_ieee_quantizer (minifloat.py:199-239)
 * and integer_quantizer (integer.py:77-95) together with block()/unblock() (utils.py:261-321).
 *
 * y has the logical shape [L, R, C]; it is written contiguous row-major, or — when
 * `transpose_out` != 0 — as [L, C, R] (used to hand V to the PV GEMM K-major).
 * fp32 results are bit-identical to the reference's torch emulation; bf16 results are
 * the same values rounded to nearest-even (exact whenever the format has <= 8
 * significant bits, except |x| <= 1e-8 pass-through elements).
 * `ws` must hold bq_quantize_workspace_bytes(...) bytes; contents need no initialisation.
 * ---------------------------------------------------------------------------------------------- */
BQ_API size_t bq_quantize_workspace_bytes(const bq_format* fmt, const bq_tensor3* x);
BQ_API int bq_quantize(const bq_format* fmt, const bq_tensor3* x_desc, const float* x, void* y, int32_t y_dtype,
                int32_t transpose_out, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * SiLU(gate) * up + quantize.  y = Q_fmt( silu(gate) * up ), the input of Llama's down_proj:
 * `self.down_proj(self.act_fn(self.gate_proj(x)) * self.up_proj(x))` (models/llama_quantized/modeling_llama.py:246) followed by
 * down_proj's x-quantizer (quantized_modules/linear.py:63-71).  gate and up share the layout `desc`; blocks [1, b] along
 * the last dim, block_fp / block_minifloat / block_log; other layouts return BQ_ERR_UNSUPPORTED (the caller then composes
 * torch silu/mul with bq_quantize).  silu(g) = g / (1 + expf(-g)) as torch-CUDA evaluates it, so fp32 results are
 * bit-identical to F.silu(gate) * up followed by the quantizer.  Workspace as for bq_quantize.
 * ---------------------------------------------------------------------------------------------- */
BQ_API int bq_silu_mul_quantize(const bq_format* fmt, const bq_tensor3* desc, const float* gate, const float* up, void* y,
                                int32_t y_dtype, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * LayerNorm / RMSNorm + quantize.  y_k = Q_fmts[k]( norm(x) ), k < n_out <= 3, each written as bf16 [rows][H].
 * Replaces nn.LayerNorm (models/opt_quantized/modeling_opt.py:386,:414) / LlamaRMSNorm
 * (models/llama_quantized/modeling_llama.py:79-92) followed by the x-quantizers (quantized_modules/linear.py:63-71) of
 * the Linears that consume the normalised tensor: q/k/v_proj read ONE tensor (modeling_opt.py:206,224-225), so it is
 * normalised once and emitted once per DISTINCT data_in format.  beta == NULL selects RMSNorm (y = gamma * x * rstd).
 * fmts: block_fp / block_minifloat with block [1,16]; H % 16 == 0, H <= 8192; x rows at stride ldx.
 * ---------------------------------------------------------------------------------------------- */
BQ_API int bq_norm_quantize(const float* x, int64_t rows, int64_t H, int64_t ldx, const float* gamma, const float* beta, float eps,
                            int32_t n_out, const bq_format* fmts, void* const* outs_bf16, void* stream);

/* Exhaustive device self-test: compares the exponent-field shortcuts the quantizers use for
 * ceil/floor/rint(log2f(x)) with libdevice's log2f on every positive finite fp32 bit pattern.
 * Writes three mismatch counters (ceil, floor, rint) to mismatches_dev3 (device memory, 3 x uint64). */
/* Rotary position embedding + the two operand quantizers of matmul_0, token-major (Llama).  Replaces apply_rotary_pos_emb_*
 * (quantized_functions/rotary_positional_encoding.py:142-167) followed by the x- / y-quantizers of the QK^T matmul
 * (models/llama_quantized/modeling_llama.py:309-314 -> quantized_functions/matmul.py:166-196):
 *   Qq[b,s,h,:] = Q_fq(q * cos[pos] + rotate_half(q) * sin[pos])   blocks of 16 along head_dim
 *   Kq[b,s,h,e] = Q_fk(k * cos[pos] + rotate_half(k) * sin[pos])   blocks of 16 consecutive positions s at fixed (h, e)  (= blocks of k^T)
 * q, k fp32 [B][S][heads*head_dim] with token strides ldq / ldk; cos / sin tables fp32 [table_rows][head_dim] ALREADY quantised by
 * the rotary table quantizer; position_ids int64 [B][S] or NULL (position = s; needs table_rows >= S).  Explicit positions must lie
 * in [0, table_rows): the kernel clamps to that range (never reads outside the tables); raising the reference's IndexError is the
 * host binding's job (rotary_positional_encoding.py does).  Outputs dense bf16 [B][S][heads*head_dim].
 * Formats: block_fp / block_minifloat, blocks [1,16]; head_dim % 32 == 0, S % 16 == 0. */
BQ_API int bq_rope_quantize(const float* q, const float* k, const float* cos_table, const float* sin_table, const int64_t* position_ids,
                            int64_t table_rows, int64_t B, int64_t S, int32_t heads, int32_t head_dim, int64_t ldq, int64_t ldk,
                            const bq_format* fq, const bq_format* fk, void* Qq_bf16, void* Kq_bf16, void* stream);
/* ------------------------------------------------------------------------------------------------
 * Attention of the formats whose matmuls keep an UNQUANTISED fp32 operand — block_log: generic_matmul_block_log quantises x only,
 * quantized_functions/matmul.py:286-297 — and of every geometry the one-kernel attention below rejects.  The S x S scores go through
 * HBM once in fp32 and once as bf16 probabilities instead of the reference's ~10 fp32 passes + two ~45-kernel quantizer calls:
 *
 *   bq_rope_quantize_split     Llama RoPE (as bq_rope_quantize) of q followed by matmul_0's x-quantizer -> Qq bf16 HEAD-major
 *                              [B][heads][S][head_dim]; RoPE of k followed by the error-free split of the fp32 result into three bf16
 *                              planes, k = k0 + k1 + k2 -> [B][3][heads][S][head_dim].  fq: block_fp / block_minifloat / block_log, [1,16].
 *                              cos_table == sin_table == NULL: no rotation (attention without rotary embedding).
 *   bq_split3_bf16_transposed  v fp32 [B][S][heads*head_dim] (token stride ldv) -> planes of v^T, [B][3][heads][head_dim][S] bf16
 *                              (keys contiguous: the K-major B operand of P @ V).  S % 2 == 0, heads*head_dim % 32 == 0.
 *   bq_bmm_split_tn            batched form of bq_gemm_split_tn for bf16 planes laid out [plane][batch][rows][K]:
 *                              C[b][m][n] = sum_t A_plane[term_a[t]][b][m][:] . B_plane[term_b[t]][b][n][:]; C row stride ldc, batch
 *                              stride sc — batch-major (sc >= M*ldc) or interleaved in the rows (sc >= N, ldc >= batch*sc: the heads of
 *                              one sequence written into the token-major activation).  With a power-of-two (block_log) or otherwise
 *                              bf16-exact x as the single A plane and the three planes of y, every product is exact: the result is the
 *                              reference's fp32 matmul up to accumulation order.
 *                              causal: 0 none; 1 (M == N, QK^T) output tiles wholly above the diagonal are neither computed nor
 *                              stored — their contents are undefined and bq_softmax_quantize(causal) never reads them; 2 (M == K,
 *                              P @ V) the K loop of a row tile stops at the last key its rows can see (P is zero beyond it).
 *   bq_softmax_quantize        P = Q_fp(softmax(max(scores / score_div + mask, finfo.min)))  in bf16 — the reference's chain
 *                              models/llama_quantized/modeling_llama.py:314-337, models/opt_quantized/modeling_opt.py:262-312,
 *                              bert_quantized/modeling_bert.py:370-435 ending in the x-quantizer of matmul_1 / bmm_1 (blocks of 16 keys).
 *                              scores fp32 [batch][Sq][lds] (batch stride ss), P bf16 [batch][Sq][ldp] (batch stride sp); batch = B*heads;
 *                              causal != 0: key j > query i masked (Sq == Sk); key_mask: optional bitmap [B][key_mask_words] (bit i of
 *                              word w = key 32w+i takes part).  causal == 2: as 1, and probabilities beyond the 256-key boundary
 *                              that follows the query are not written (the causal == 2 bq_bmm_split_tn never reads them; a fully
 *                              masked row is still written in full).  Masked scores are finfo.min exactly like the reference's additive masks
 *                              leave them, so a fully masked row comes out uniform over ALL keys as it does there.  expf / IEEE division
 *                              like torch's softmax; the row sum is accumulated in a different order.  Sk % 16 == 0.
 *                              fp: block_fp / block_minifloat / block_log, blocks [1,16].
 * block_log in a 16-bit carrier (this function, bq_rope_quantize_split, bq_norm_quantize, the GEMM epilogue quantizer): block-local
 * "carrier rule" — an all-zero block stays 0 (the reference fills it with 2^(ceil(log2 g) - 127), g = the tensor's smallest non-zero
 * block maximum) and outputs the reference puts below 2^-126 are 0 or 2^-126; every other output is bit-identical.
 * ---------------------------------------------------------------------------------------------- */
BQ_API int bq_rope_quantize_split(const float* q, const float* k, const float* cos_table, const float* sin_table,
                                  const int64_t* position_ids, int64_t table_rows, int64_t B, int64_t S, int32_t heads, int32_t head_dim,
                                  int64_t ldq, int64_t ldk, const bq_format* fq, void* Qq_bf16, void* K_planes_bf16, void* stream);
BQ_API int bq_split3_bf16_transposed(const float* v, void* planes_bf16, int64_t B, int64_t S, int32_t heads, int32_t head_dim, int64_t ldv,
                                     void* stream);
BQ_API int bq_bmm_split_tn(const void* A_planes, const void* B_planes, float* C, int64_t batch, int64_t M, int64_t N, int64_t K,
                           int32_t planes_a, int32_t planes_b, int32_t n_terms, const int32_t* term_a, const int32_t* term_b,
                           int64_t ldc, int64_t sc, int32_t causal, void* stream);
/* A/B switch of bq_softmax_quantize: 1 (default) = row staged in shared memory, rolled loops; 0 = row held in registers (v1).  Same results. */
BQ_API void bq_set_softmax_smem_rows(int on);
BQ_API int bq_softmax_quantize(const bq_format* fp, const float* scores, void* P_bf16, int64_t batch, int64_t heads, int64_t Sq, int64_t Sk,
                               int64_t lds, int64_t ss, int64_t ldp, int64_t sp, float score_div, int32_t causal, const uint32_t* key_mask,
                               int64_t key_mask_words, void* stream);
BQ_API int bq_selftest_log2(unsigned long long* mismatches_dev3, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Dense bf16 GEMM on the 5th-gen tensor cores (TMA -> smem -> tcgen05.mma -> TMEM -> fp32).
 *   C[b][m][n] = sum_k A[b][m][k] * B[b][n][k]  (+ bias[n])          ("TN": both operands K-major)
 * A: bf16 [batch][M][K] (row stride lda, batch stride sa), B: bf16 [batch][N][K] (ldb, sb; sb = 0
 * broadcasts one weight matrix), C: fp32 [batch][M][N] (ldc, sc).  K-major operands must have
 * lda, ldb multiples of 8 elements and 16-byte aligned bases.
 * Replaces the fp32 F.linear / torch.matmul / torch.bmm calls of quantized_modules/linear.py:62,71,76
 * and quantized_functions/matmul.py:25,196 for operands that are bf16-exact.
 * ---------------------------------------------------------------------------------------------- */
BQ_API int bq_gemm_bf16_tn(const void* A, const void* B, float* C, const float* bias, int64_t batch, int64_t M, int64_t N,
                    int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int64_t sa, int64_t sb, int64_t sc,
                    void* stream);

/* ------------------------------------------------------------------------------------------------
 * Same GEMM (batch 1) with a fused epilogue, for the glue BETWEEN two quantized ops of a decoder layer:
 *   v = acc + bias[n];  v *= scale;  v = act(v);  v = residual[m][n] + v;  v = Q_qfmt(v);  C[m][n] = v (fp32 or bf16)
 * in exactly this order — the order of the reference's separate torch ops:
 *   F.linear bias add (quantized_modules/linear.py:71), `q_proj(x) * self.scaling` (models/opt_quantized/modeling_opt.py:206),
 *   ReLU between fc1 and fc2 (:416-418), `residual + hidden_states` (:395, :424), and the x-quantizer of the op that
 *   consumes the result (linear.py:63-71; bmm_0 / bmm_1 operands, quantized_functions/matmul.py:165-193).
 * qfmt (optional): block_fp or block_minifloat with blocks of 16 — qdir 0: along N (16 consecutive output features:
 * data_in of a following Linear, q and v of the attention bmms), qdir 1: along M (16 consecutive tokens at one feature:
 * the k^T operand of bmm_0, modeling_opt.py:246; needs M % 16 == 0).  With qfmt the natural out_dtype is BQ_BF16 (exact
 * carrier); fp32 output of quantised values is also allowed.  N % 32 == 0; bias / residual / C 16-byte aligned.
 * act 2 — gated SiLU, the Llama MLP's `down_proj(act_fn(gate_proj(x)) * up_proj(x))` (models/llama_quantized/modeling_llama.py:84)
 * in one launch: B holds BOTH quantised weights interleaved in groups of 16 rows ([gate f..f+16), [up f..f+16), ...; N = 2 * features,
 * bias — if any — interleaved the same way), the epilogue forms silu(gate) * up = gate / (1 + expf(-gate)) * up (torch-CUDA's op
 * order), applies qfmt (required; qdir 0 — the x-quantizer of down_proj, block_fp / block_minifloat / block_log carrier rule) and
 * stores bf16 C[M][N / 2] (ldc >= N / 2).  No residual, scale 1, no replicas.  Same bits as two GEMMs + bq_silu_mul_quantize.
 * ---------------------------------------------------------------------------------------------- */
typedef struct bq_gemm_epilogue {
  const float* bias;       /* [N] or NULL                                   */
  const float* residual;   /* fp32 [M][ldr] or NULL                         */
  int64_t ldr;
  float scale;             /* 1.0f = none                                   */
  int32_t act;             /* 0 none, 1 ReLU, 2 gated SiLU (see above)      */
  int32_t out_dtype;       /* bq_dtype of C                                 */
  const bq_format* qfmt;   /* NULL = no quantisation                        */
  int32_t qdir;            /* 0: blocks along N, 1: blocks along M          */
  /* Fused all-gather (column-parallel Linear): every output tile is ALSO stored to the same [m][n] position of
   * n_replicas further buffers with the same ldc and dtype as C — peer GPUs' copies of the gathered output, mapped with
   * bq_ipc_import, written over NVLink from the epilogue while the other tiles are still being multiplied.  The caller
   * passes pointers already offset to this rank's first column and orders visibility with bq_peer_barrier. */
  int32_t n_replicas;      /* 0..BQ_MAX_REPLICAS                            */
  void* replicas[7];
} bq_gemm_epilogue;
#define BQ_MAX_REPLICAS 7
BQ_API int bq_gemm_bf16_tn_ex(const void* A, const void* B, void* C, const bq_gemm_epilogue* ep, int64_t M, int64_t N, int64_t K,
                              int64_t lda, int64_t ldb, int64_t ldc, void* stream);

/* ------------------------------------------------------------------------------------------------
 * q_proj / k_proj of a Llama layer with the rotary position embedding AND matmul_0's operand quantizer in the GEMM epilogue:
 *   y = A @ B^T (+ bias);  y = rope(y, cos[pos], sin[pos]) per head of head_dim features;  C = Q_qfmt(y) as bf16 [M][ldc]
 * replacing (models/llama_quantized/modeling_llama.py:274-276, :309-314; quantized_functions/rotary_positional_encoding.py:27-36;
 * quantized_functions/matmul.py:165-193) an fp32 GEMM output, ~12 element-wise kernels and a quantizer call.  Tables: fp32
 * [table_rows][head_dim], already quantised by the caller exactly as the reference quantises them; position of row m =
 * position_ids[m] (clamped into the table) or m % S when position_ids is NULL.  qdir 0: blocks of 16 along the features (q, the x
 * operand of matmul_0); qdir 1: blocks of 16 consecutive tokens at one feature (k^T, its y operand; M % 16 == 0, S % 16 == 0).
 * head_dim 64 or 128, N % head_dim == 0; block_fp / block_minifloat.  Same arithmetic and order as bq_rope_quantize: same bits.
 * ---------------------------------------------------------------------------------------------- */
BQ_API int bq_gemm_bf16_tn_rope(const void* A, const void* B, void* C_bf16, const float* bias, const bq_format* qfmt, int32_t qdir,
                                const float* cos_table, const float* sin_table, const int64_t* position_ids, int64_t table_rows,
                                int64_t S, int64_t head_dim, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc,
                                void* stream);

/* q_proj | k_proj | v_proj of one Llama layer in ONE launch: B_qkv = the three quantised weights concatenated along N ([3 * H][K]; all
 * three read the same x operand, i.e. their x-quantizers coincide), bias likewise ([3 * H] or NULL).  Segment 0 -> Cq: RoPE + fq in blocks
 * of 16 features; segment 1 -> Ck: RoPE + fk in blocks of 16 consecutive tokens (the k^T operand of matmul_0); segment 2 -> Cv: fv in
 * blocks of 16 features, no rotation (the y operand of matmul_1).  Each output bf16 [M][ldc].  H % 256 == 0 (a tile never straddles two
 * projections), M % 16 == 0, S % 16 == 0.  Same bits as three bq_gemm_bf16_tn_rope / bq_gemm_bf16_tn_ex calls; 768 tiles in one
 * persistent launch fill the last wave better than 3 x 256 at the Llama-7B shape and expose one epilogue tail instead of three. */
BQ_API int bq_gemm_bf16_tn_qkv_rope(const void* A, const void* B_qkv, void* Cq_bf16, void* Ck_bf16, void* Cv_bf16, const float* bias,
                                    const bq_format* fq, const bq_format* fk, const bq_format* fv, const float* cos_table,
                                    const float* sin_table, const int64_t* position_ids, int64_t table_rows, int64_t S,
                                    int64_t head_dim, int64_t M, int64_t H, int64_t K, int64_t lda, int64_t ldb, int64_t ldc,
                                    void* stream);

/* ------------------------------------------------------------------------------------------------
 * fp32-equivalent GEMM on the bf16 tensor cores, for the matmuls the reference leaves UNQUANTISED in fp32
 * (lm_head: models/opt_quantized/modeling_opt.py:942-944, models/llama_quantized/modeling_llama.py:772; bypass
 * layers: quantized_modules/linear.py:60-62; the y operand of block_log matmuls: quantized_functions/matmul.py:293-296).
 *   bq_split3_bf16: x (fp32, n elements, n % 4 == 0) -> three bf16 planes [3][n] with x = p0 + p1 + p2 (+- 2^-25 |x|).
 *   bq_gemm_split_tn: C[m][n] = sum_t  A_plane[term_a[t]][m][:] . B_plane[term_b[t]][n][:]   (+ bias[n]),
 *   all terms accumulated in one fp32 TMEM accumulator.  Planes are dense [planes][rows][K]; an exactly
 *   bf16-representable operand is passed as a single plane.
 * ---------------------------------------------------------------------------------------------- */
BQ_API int bq_split3_bf16(const float* x, void* planes_bf16, int64_t n, void* stream);
/* Cheaper variant with the same purpose (half the tensor work): per-row scaled fp16 planes.
 *   bq_split2_f16_rows: x fp32 [rows][K] (row stride ldx) -> fp16 planes [2][rows][K] with x * 2^e(row) = hi + lo up to 2^-22 of
 *   the row max, and inv_scale[row] = 2^-e(row).
 *   bq_gemm_split16_tn: C[m][n] = a_inv_scale[m] * b_inv_scale[n] * sum_t A_plane[term_a[t]][m][:] . B_plane[term_b[t]][n][:] (+ bias[n]);
 *   the three terms (lo,hi), (hi,lo), (hi,hi) give ~2^-21 relative error per product. */
BQ_API int bq_split2_f16_rows(const float* x, int64_t rows, int64_t K, int64_t ldx, void* planes_f16, float* inv_scale, void* stream);
BQ_API int bq_gemm_split16_tn(const void* A_planes_f16, const void* B_planes_f16, float* C, const float* bias,
                              const float* a_inv_scale, const float* b_inv_scale, int64_t M, int64_t N, int64_t K,
                              int32_t n_terms, const int32_t* term_a, const int32_t* term_b, int64_t ldc, void* stream);
/* Batched form for matmul / bmm whose operands are not bf16-exact — generic_matmul_block_log leaves y unquantised in fp32
 * (quantized_functions/matmul.py:293-296), formats wider than 8 significant bits — instead of an fp32 SIMT library GEMM:
 *   C[b][m][n] = a_inv_scale[b*M+m] * b_inv_scale[b*N+n] * sum_t A_plane[term_a[t]][b][m][:] . B_plane[term_b[t]][b][n][:]
 * planes are fp16 [2][batch][rows][K] as written by ONE bq_split2_f16_rows call over batch*rows rows; C row stride ldc, batch
 * stride sc (elements). */
BQ_API int bq_bmm_split16_tn(const void* A_planes_f16, const void* B_planes_f16, float* C, const float* a_inv_scale,
                             const float* b_inv_scale, int64_t batch, int64_t M, int64_t N, int64_t K, int32_t n_terms,
                             const int32_t* term_a, const int32_t* term_b, int64_t ldc, int64_t sc, void* stream);
BQ_API int bq_gemm_split_tn(const void* A_planes, const void* B_planes, float* C, const float* bias, int64_t M, int64_t N,
                            int64_t K, int32_t planes_a, int32_t planes_b, int32_t n_terms, const int32_t* term_a,
                            const int32_t* term_b, int64_t ldc, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Quantized Linear, PTQ steady state.  Replaces _LinearBase.forward (quantized_modules/linear.py:59-76):
 *   y[M,N] = F.linear(Q_fx(x[M,K]), Wq, bias_q)
 * Wq is the weight already quantised by bq_quantize to bf16 [N][K] (the in-place PTQ overwrite of
 * linear.py:66-70, held as a bf16 cache), bias_q the quantised fp32 bias or NULL.
 * ---------------------------------------------------------------------------------------------- */
BQ_API size_t bq_linear_workspace_bytes(const bq_format* fx, int64_t M, int64_t K);
BQ_API int bq_linear(const bq_format* fx, const float* x, int64_t M, int64_t K, int64_t ldx, const void* Wq_bf16, int64_t N,
              const float* bias_q, float* y, int64_t ldy, void* ws, size_t ws_bytes, void* stream);
/* Same contract in ONE launch: the x-quantizer runs in the GEMM PROLOGUE — raw fp32 A tiles arrive by TMA, transform warps
 * quantise them in shared memory (same arithmetic as bq_quantize) and hand tcgen05.mma a bf16 K-major tile; no bf16 copy of x
 * in HBM, no workspace.  Each N-tile of a row block re-quantises the same A tile, which is why bq_linear (two launches) stays
 * the default for wide N (measured A/B: DESIGN.md §3).  fx: block_fp / block_minifloat, blocks [1,16]; K % 64 == 0, N % 32 == 0. */
BQ_API int bq_linear_fused(const bq_format* fx, const float* x, int64_t M, int64_t K, int64_t ldx, const void* Wq_bf16, int64_t N,
                           const float* bias_q, float* y, int64_t ldy, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Packed weight cache.  The reference's PTQ state is the quantised weight (quantized_modules/linear.py:66-70) and its cost model
 * charges `width` bits per element plus one shared exponent per block (quantized_layer_profiler.py:18-27): w + 0.5 bits/element
 * for block_fp, blocks of 16.  bq_pack_weight stores exactly that: per row, per group of 256 K-elements, 32*w bytes of fields
 * (16 per block, w bits each: sign in the top bit, magnitude below, little-endian bit order) followed by 16 exponent bytes
 * (E + exponent_bias, one per block; value = (-1)^sign * magnitude * 2^(E - (w-1))).  Wq: fp32 [N][ldw] ALREADY on the block_fp
 * grid (after the PTQ overwrite); *mismatches (device) counts elements the packed form does not reproduce bit for bit — the
 * reference's pass-through elements (|x| <= 1e-8 left unquantised, block_fp.py:93-94), rounded to the grid here.
 * bq_gemm_packed_tn: y = A_bf16[M][K] @ unpack(packed)[N][K]^T (+ bias), weights decoded to bf16 inside the mainloop (exact:
 * <= 7 magnitude bits x a power of two).  K % 256 == 0, N % 32 == 0, block_fp with 2 <= width <= 8.
 * ---------------------------------------------------------------------------------------------- */
BQ_API size_t bq_packed_weight_bytes(const bq_format* fw, int64_t N, int64_t K);
BQ_API int bq_pack_weight(const bq_format* fw, const float* Wq, int64_t N, int64_t K, int64_t ldw, void* packed,
                          unsigned long long* mismatches, void* stream);
BQ_API int bq_gemm_packed_tn(const void* A_bf16, const void* packed, const bq_format* fw, float* y, const float* bias, int64_t M,
                             int64_t N, int64_t K, int64_t lda, int64_t ldy, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Quantized bmm / matmul.  Replaces generic_matmul_block_fp / _block_minifloat / _block_log /
 * _minifloat_denorm (quantized_functions/matmul.py:146-297, :46-78):
 *   out[b] = Q_fx(x[b]) @ Q_fy(y[b]),  x: [batch, M, K] fp32 contiguous (blocks along K),
 *   y: logical [batch, K, N] fp32 (blocks along N) given with element strides (syK, syN) so that
 *   k^T views (syK = 1) and plain [K,N] tensors (syN = 1) both work without a copy.
 * fy == NULL or kind NONE leaves y unquantised (block_log quirk, matmul.py:293-296) — y is then
 * rounded to bf16, which is NOT exact; the Python mirror routes that case elsewhere.
 * ---------------------------------------------------------------------------------------------- */
BQ_API size_t bq_bmm_workspace_bytes(const bq_format* fx, const bq_format* fy, int64_t batch, int64_t M, int64_t K, int64_t N);
BQ_API int bq_bmm(const bq_format* fx, const bq_format* fy, const float* x, const float* y, int64_t batch, int64_t M,
           int64_t K, int64_t N, int64_t sy_batch, int64_t syK, int64_t syN, float* out, void* ws, size_t ws_bytes,
           void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused causal quantized attention.  Replaces bmm_0 -> (+causal mask, max finfo.min) -> softmax -> bmm_1 of
 * models/opt_quantized/modeling_opt.py:246-312 (and matmul_0 / sqrt(d) -> ... -> matmul_1 of
 * models/llama_quantized/modeling_llama.py:309-344) when the mask is purely causal:
 *   out[b,s,h,:] = Q_fp( softmax_row( (Qq[b,:,h,:] Kq[b,:,h,:]^T) / score_div , causal ) ) @ Vq[b,:,h,:]
 * Qq / Kq / Vq: bf16 [B,S,H,d] operands ALREADY quantised (bq_quantize or a quantising GEMM epilogue; q: data_in of
 * bmm_0, blocks along d; k: weight of bmm_0, blocks along S; v: weight of bmm_1, blocks along d), token strides
 * ldq/ldk/ldv (elements).  fp: format of the probabilities (data_in of bmm_1), block [1,16], block_fp or block_minifloat.
 * out: fp32 [B,S,H,d] with token stride ldo.  d must be 64 or 128.  Scores and probabilities never touch HBM.
 * score_div: scores are MULTIPLIED by 1.0f/score_div — what torch-CUDA evaluates for `tensor / python_float`.
 *
 * bq_attention_causal_q additionally applies the x-quantizer of the Linear that consumes the attention output
 * (out_proj / o_proj, quantized_modules/linear.py:63-71; format fo, blocks [1,16] along the hidden dim) in the
 * epilogue and writes the exact quantised values as bf16 [B,S,H,d] — the A operand of bq_gemm_bf16_tn.
 * ---------------------------------------------------------------------------------------------- */
BQ_API int bq_attention_causal(const bq_format* fp, const void* Qq, const void* Kq, const void* Vq, float* out, int64_t B,
                               int64_t H, int64_t S, int64_t d, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo,
                               float score_div, void* stream);
/* 1: numerators of the softmax use libdevice expf (bit-identical to torch's exp(x - max)); 0 (default): ex2.approx of a fused
 * multiply-add argument, ~|x - max| * 1.44 ulp less accurate, 30 % fewer instructions per score. */
BQ_API void bq_set_attention_precise_exp(int on);
BQ_API int bq_get_attention_precise_exp(void);           /* current mode, so that a benchmark can state which one it timed */
/* head_dim 64 only.  1 (default): two independent softmax pipelines per CTA with the quantised probabilities written to tensor
 * memory (tcgen05.st) and PV issued with its A operand read from TMEM; 0: the single-pipeline kernel that stages P through shared
 * memory (what head_dim 128 uses).  Same arithmetic, same results; A/B measurement switch. */
BQ_API void bq_set_attention_dual_pipeline(int on);
BQ_API int bq_get_attention_dual_pipeline(void);
/* General form: causal != 0 -> decoder mask (keys above the diagonal excluded); causal == 0 -> bidirectional (BERT,
 * bert_quantized/modeling_bert.py:366-435).  key_mask: optional bitmap [B][key_mask_words] (bit i of word w set = key 32*w + i
 * takes part; the reference adds finfo.min to the other scores — opt_quantized/modeling_opt.py:520-548 — whose probabilities are
 * exactly 0).  key_mask_words >= 4 * ceil(S / 128), bits of keys >= S clear; REQUIRED when causal == 0.  Every query row must keep at
 * least one key (the reference's all-masked rows degenerate to a uniform distribution over ALL keys; callers route such batches to
 * the op-by-op path).  fo NULL: fp32 output (out = float*), else bf16 output quantised for the consuming Linear. */
BQ_API int bq_attention_masked(const bq_format* fp, const bq_format* fo, const void* Qq, const void* Kq, const void* Vq, void* out,
                               int64_t B, int64_t H, int64_t S, int64_t d, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo,
                               float score_div, int32_t causal, const uint32_t* key_mask, int64_t key_mask_words, void* stream);
BQ_API int bq_attention_causal_q(const bq_format* fp, const bq_format* fo, const void* Qq, const void* Kq, const void* Vq,
                                 void* out_bf16, int64_t B, int64_t H, int64_t S, int64_t d, int64_t ldq, int64_t ldk,
                                 int64_t ldv, int64_t ldo, float score_div, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Token cross-entropy (perplexity numerator).  Replaces the loss tail of the reference's causal-LM forward,
 *     shift_logits = logits[..., :-1, :].contiguous(); shift_labels = labels[..., 1:].contiguous()
 *     loss = CrossEntropyLoss()(shift_logits.view(-1, V), shift_labels.view(-1))
 * (models/opt_quantized/modeling_opt.py:1086-1098, models/llama_quantized/modeling_llama.py:867-879; the number
 * eval/eval_lm.py:41-63 turns into perplexity) with ONE streaming read of the logits — no shifted copy, no
 * log-softmax tensor.  logits: fp32 [n_seq, seq_len, vocab], row stride ld; labels: int64 [n_seq, seq_len] contiguous.
 * shift = 1: row (b, t), t < seq_len - 1, is scored against labels[b, t + 1]; shift = 0: against labels[b, t].
 * Rows whose target equals ignore_index are skipped (mean over the others, like reduction="mean"); targets outside
 * [0, vocab) — a device assert in torch — are skipped as well.  out2[0] = mean loss (NaN when no row is valid),
 * out2[1] = number of valid rows.  ws: bq_token_ce_workspace_bytes(n_seq, seq_len) bytes (one fp32 per row).
 * Deterministic (fixed reduction order).  Log-sum-exp differs from torch's by <= ~1e-6 absolute (ex2.approx).
 * ---------------------------------------------------------------------------------------------- */
BQ_API size_t bq_token_ce_workspace_bytes(int64_t n_seq, int64_t seq_len);
BQ_API int bq_token_ce_mean(const float* logits, int64_t n_seq, int64_t seq_len, int64_t vocab, int64_t ld, const int64_t* labels,
                            int32_t shift, int64_t ignore_index, float* out2, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Launch accounting (measurement support for bench.py; no reference counterpart).
 * Every kernel this library launches is counted per kernel id.  With profiling enabled the library
 * additionally brackets each launch with CUDA events on the launching stream; bq_profile_read
 * synchronises those events and returns the summed device time.
 * ---------------------------------------------------------------------------------------------- */
BQ_API void bq_set_stream_quantizer(int on);           /* 0: force the per-slot quant_rows_kernel instead of the bulk-copy streaming kernel (A/B measurement) */
BQ_API void bq_set_norm_warp_rows(int on);             /* 0: force the row-per-CTA norm_quant_kernel for H <= 2048 too (A/B measurement, tests) */
BQ_API void bq_set_cta_pairs(int on);                  /* 0: force cta_group::1 GEMM tiles (A/B measurement) */
BQ_API void bq_set_small_tiles(int on);                /* 0: always the largest GEMM tile the shape admits; 1 (default): 128 x 128 tiles when the
                                                          problem would not fill the chip with larger ones (A/B measurement) */
BQ_API void bq_set_pdl(int on);                        /* 1: the tcgen05 GEMM, attention and norm+quantize kernels are launched with
                                                          programmatic stream serialization — a kernel's set-up (barriers, TMEM, descriptors)
                                                          overlaps the tail of its predecessor, every global access waits for the predecessor's
                                                          completion (griddepcontrol.wait); 0 (default — measured 0.7 % slower with it on the
                                                          power-capped headline step): plain stream order */
BQ_API int bq_get_pdl(void);
BQ_API int bq_kernel_count(void);
BQ_API const char* bq_kernel_name(int kernel_id);
BQ_API int64_t bq_launch_count(int kernel_id);          /* launches since load (kernel_id < 0: all kernels) */
BQ_API void bq_profile_enable(int on);                  /* turning it on clears earlier records */
BQ_API int bq_profile_read(int kernel_id, double* total_ms, int64_t* launches);

/* ------------------------------------------------------------------------------------------------
 * Peer memory over NVLink / NVSwitch — the one exchange step of the path (column-parallel Linear).
 * The reference has no counterpart (single device).  One process per GPU; buffers stay owned by the caller (torch):
 *   bq_ipc_export   handle of the cudaMalloc allocation that contains dev_ptr, plus dev_ptr's offset inside it
 *   bq_ipc_import   maps a peer process's allocation (peer access enabled lazily); *base = start of the mapping,
 *                   *ptr = base + offset.  One mapping per (process, allocation): callers cache it.
 *   bq_ipc_release  unmaps (pass *base).
 *   bq_peer_barrier stream-ordered barrier between the `world` processes of a job, on flags that live in peer-mapped
 *                   memory: signals[r] points at rank r's flag block (BQ_PEER_FLAG_WORDS x uint32, zeroed once);
 *                   rank `rank` release-increments word [rank] of every peer's block and acquire-waits until its own
 *                   block shows `epoch` arrivals from every peer (epoch = number of barriers so far, counted by the
 *                   caller identically on all ranks).  All writes issued by earlier work on `stream` (including
 *                   epilogue stores into peers) are visible to the peers' later work.  A wait longer than
 *                   timeout_ms sets word [BQ_PEER_FLAG_TIMEOUT] of the own block and returns (never hangs the GPU).
 *                   epoch == 0: device-resident count (word [BQ_PEER_FLAG_EPOCH] of the own block, advanced by the kernel) — no
 *                   per-call host state, so the launch can be captured in a CUDA graph and replayed.  Use one mode per flag block.
 * ---------------------------------------------------------------------------------------------- */
typedef struct bq_ipc_handle {
  unsigned char reserved[64];   /* cudaIpcMemHandle_t */
  int64_t offset;               /* of dev_ptr inside the allocation */
  int64_t size;                 /* of the allocation */
} bq_ipc_handle;
#define BQ_PEER_FLAG_WORDS 64
#define BQ_PEER_FLAG_TIMEOUT 32
#define BQ_PEER_FLAG_EPOCH 33      /* epoch == 0 passed to bq_peer_barrier[_ex]: the barrier count is kept (and advanced) here on the device */
BQ_API int bq_ipc_export(const void* dev_ptr, bq_ipc_handle* out);
BQ_API int bq_ipc_import(const bq_ipc_handle* h, void** base, void** ptr);
BQ_API int bq_ipc_release(void* base);
BQ_API int bq_peer_barrier(void* const* signals, int32_t rank, int32_t world, uint32_t epoch, int32_t timeout_ms, void* stream);
/* Same; a timed-out wait additionally writes 1 to *host_error_flag — a uint32 in MAPPED PINNED HOST memory (device-accessible under
 * unified addressing), so that the host notices a stalled peer on its next call without synchronising the device. */
BQ_API int bq_peer_barrier_ex(void* const* signals, int32_t rank, int32_t world, uint32_t epoch, int32_t timeout_ms,
                              uint32_t* host_error_flag, void* stream);
/* Copies a strided slab [rows][row_bytes] (16-byte multiples, 16-byte aligned) from local memory to the SAME position of n_dst
 * (<= BQ_MAX_REPLICAS) peer-mapped buffers: the gather of a slab that no GEMM epilogue produced (this rank's heads of the
 * attention output in the tensor-parallel decoder layer).  Order visibility with bq_peer_barrier. */
BQ_API int bq_peer_push(const void* src, void* const* dst, int32_t n_dst, int64_t rows, int64_t row_bytes, int64_t src_stride_bytes,
                        int64_t dst_stride_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BQ_B200_H */
