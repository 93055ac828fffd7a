// ops.cu — the consumers of the quantizers: quantized Linear and quantized bmm/matmul (C ABI).
//
//   bq_linear  replaces _LinearBase.forward, PTQ steady state (quantized_modules/linear.py:59-76):
//              y = F.linear(Qx(x), Wq, bq).  Wq is the bf16 cache of the quantised weight.
//   bq_bmm     replaces generic_matmul_* (quantized_functions/matmul.py:146-297):
//              out[b] = Qx(x[b]) @ Qy(y[b]).
//
// Round-1 structure: activation quantisation writes bf16 (6 B/element instead of the reference's
// fp32 round trip) into a caller-owned workspace, then the tcgen05 GEMM consumes it.
#include "bq_internal.h"

namespace bq {

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static bool bf16_exact(const bq_format* f) {
  switch (f->kind) {
    case BQ_KIND_BLOCK_FP: return f->width - 1 <= 8;
    case BQ_KIND_BLOCK_MINIFLOAT:
    case BQ_KIND_MINIFLOAT_IEEE: return f->width - f->exponent_width - 1 + 1 <= 8;
    case BQ_KIND_MINIFLOAT_DENORM: return f->width - f->exponent_width - 1 <= 8;
    case BQ_KIND_BLOCK_LOG: return true;
    case BQ_KIND_INTEGER: return f->width - 1 <= 8;
    default: return false;   // NONE: fp32 passthrough is not bf16 exact
  }
}

}  // namespace bq

extern "C" {

size_t bq_linear_workspace_bytes(const bq_format* fx, int64_t M, int64_t K) {
  if (!fx || M < 0 || K < 0) return 0;
  bq_tensor3 t = {1, M, K, M * K, K, 1};
  int64_t Kp = (K + 7) / 8 * 8;
  return bq::align_up((size_t)M * Kp * 2, 256) + bq::quantize_ws_bytes(fx, &t) + 256;
}

int bq_linear(const bq_format* fx, const float* x, int64_t M, int64_t K, int64_t ldx, const void* Wq_bf16, int64_t N,
              const float* bias_q, float* y, int64_t ldy, void* ws, size_t ws_bytes, void* stream) {
  if (!fx || M < 0 || K < 0 || N < 0) return BQ_ERR_BAD_ARG;
  if (M == 0 || N == 0) return BQ_OK;
  if (!x || !Wq_bf16 || !y || !ws) return BQ_ERR_BAD_ARG;
  if (K % 8) return BQ_ERR_UNSUPPORTED;            // TMA needs 16-byte row pitch for the bf16 operands
  if (!bq::bf16_exact(fx)) return BQ_ERR_NOT_BF16_EXACT;
  if (ws_bytes < bq_linear_workspace_bytes(fx, M, K) || ((uintptr_t)ws % 256)) return BQ_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* xq = (uint8_t*)ws;
  size_t xq_bytes = bq::align_up((size_t)M * K * 2, 256);
  bq_tensor3 t = {1, M, K, M * ldx, ldx, 1};
  bq_format f = *fx;
  f.fold_zero = 0;                                 // sign of zero is irrelevant to the product
  int rc = bq::quantize_impl(&f, &t, x, xq, BQ_BF16, 0, xq + xq_bytes, ws_bytes - xq_bytes, st);
  if (rc) return rc;
  return bq::gemm_bf16_tn_impl(xq, Wq_bf16, y, bias_q, 1, M, N, K, K, K, ldy, 0, 0, 0, st);
}

size_t bq_bmm_workspace_bytes(const bq_format* fx, const bq_format* fy, int64_t batch, int64_t M, int64_t K, int64_t N) {
  if (!fx || batch < 0 || M < 0 || K < 0 || N < 0) return 0;
  bq_tensor3 tx = {batch, M, K, M * K, K, 1};
  bq_tensor3 ty = {batch, K, N, K * N, N, 1};
  size_t w = bq::align_up((size_t)batch * M * K * 2, 256) + bq::align_up((size_t)batch * N * K * 2, 256);
  size_t q = bq::quantize_ws_bytes(fx, &tx);
  if (fy) q = std::max(q, bq::quantize_ws_bytes(fy, &ty));
  return w + q + 256;
}

int bq_bmm(const bq_format* fx, const bq_format* fy, const float* x, const float* y, int64_t batch, int64_t M,
           int64_t K, int64_t N, int64_t sy_batch, int64_t syK, int64_t syN, float* out, void* ws, size_t ws_bytes,
           void* stream) {
  if (!fx || batch < 0 || M < 0 || K < 0 || N < 0) return BQ_ERR_BAD_ARG;
  if (batch == 0 || M == 0 || N == 0) return BQ_OK;
  if (!x || !y || !out || !ws) return BQ_ERR_BAD_ARG;
  if (K % 8) return BQ_ERR_UNSUPPORTED;
  if (!bq::bf16_exact(fx)) return BQ_ERR_NOT_BF16_EXACT;
  bq_format none = {BQ_KIND_NONE, 0, 0, 0, 0, 1, 1, 0};
  const bq_format* fyy = fy ? fy : &none;
  if (fyy->kind != BQ_KIND_NONE && !bq::bf16_exact(fyy)) return BQ_ERR_NOT_BF16_EXACT;
  if (ws_bytes < bq_bmm_workspace_bytes(fx, fy, batch, M, K, N) || ((uintptr_t)ws % 256)) return BQ_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* xq = (uint8_t*)ws;
  size_t xq_bytes = bq::align_up((size_t)batch * M * K * 2, 256);
  uint8_t* yq = xq + xq_bytes;
  size_t yq_bytes = bq::align_up((size_t)batch * N * K * 2, 256);
  uint8_t* qws = yq + yq_bytes;
  size_t qws_bytes = ws_bytes - xq_bytes - yq_bytes;
  bq_format f = *fx;
  f.fold_zero = 0;
  bq_tensor3 tx = {batch, M, K, M * K, K, 1};
  int rc = bq::quantize_impl(&f, &tx, x, xq, BQ_BF16, 0, qws, qws_bytes, st);
  if (rc) return rc;
  // y: logical [batch, K, N], blocks along N; written transposed -> bf16 [batch][N][K] (K-major B operand)
  f = *fyy;
  f.fold_zero = 0;
  bq_tensor3 ty = {batch, K, N, sy_batch, syK, syN};
  rc = bq::quantize_impl(&f, &ty, y, yq, BQ_BF16, 1, qws, qws_bytes, st);
  if (rc) return rc;
  return bq::gemm_bf16_tn_impl(xq, yq, out, nullptr, batch, M, N, K, K, K, N, M * K, N * K, M * N, st);
}
}
