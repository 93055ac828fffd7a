// common.cu — status strings and error bookkeeping of the C ABI (include/bq.h).
#include "bq_internal.h"

namespace bq {
static thread_local char g_last_err[512] = "";
void set_last_cuda_error(const char* what, const char* file, int line) {
  snprintf(g_last_err, sizeof(g_last_err), "%s (%s:%d)", what ? what : "?", file, line);
}
}  // namespace bq

extern "C" {
const char* bq_strerror(int status) {
  switch (status) {
    case BQ_OK: return "ok";
    case BQ_ERR_BAD_ARG: return "bad argument (null/misaligned pointer or negative size)";
    case BQ_ERR_UNSUPPORTED: return "unsupported configuration";
    case BQ_ERR_BAD_FORMAT: return "bad quantisation format (width / exponent width out of range)";
    case BQ_ERR_WORKSPACE: return "workspace missing or too small";
    case BQ_ERR_CUDA: return "CUDA error";
    case BQ_ERR_NOT_BF16_EXACT: return "quantised operand is not exactly representable in bf16";
    default: return "unknown status";
  }
}
int bq_abi_version(void) { return BQ_ABI_VERSION; }
const char* bq_last_cuda_error(void) { return bq::g_last_err; }
}
