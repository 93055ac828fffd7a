// common.cu — status strings and error bookkeeping of the C ABI (include/bq.h).
#include "bq_internal.h"

#include <atomic>
#include <mutex>
#include <vector>

namespace bq {
static thread_local char g_last_err[512] = "";
void set_last_cuda_error(const char* what, const char* file, int line) {
  snprintf(g_last_err, sizeof(g_last_err), "%s (%s:%d)", what ? what : "?", file, line);
}

// ---------------------------------------------------------------------------------------------
// launch accounting
// ---------------------------------------------------------------------------------------------
int current_device() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return 0;
  return d;
}
static const char* kKernelNames[kKernCount] = {"quant_rows_kernel", "blocklog_fixup_kernel", "quant_tile_kernel",
                                               "generic_blockmax_kernel", "generic_gmin_kernel", "generic_quant_kernel",
                                               "gemm_bf16_tn_kernel", "attention_causal_kernel", "split3_kernel",
                                               "gemm_bf16_tn_kernel<epilogue>", "gemm_bf16_tn_kernel<split>", "layernorm_quant_kernel", "quant_stream_kernel",
                                               "silu_mul_quant_kernel", "ce_rows_kernel", "ce_mean_kernel", "peer_barrier_kernel", "rope_quant_kernel", "peer_push_kernel", "gemm_xform_kernel<quantize A>", "gemm_xform_kernel<packed B>",
                                               "pack_weight_kernel", "softmax_quant_kernel", "rope_split_kernel", "split3_transposed_kernel"};
static std::atomic<int64_t> g_launches[kKernCount];
static std::atomic<int> g_profiling{0};
struct EventPair { int id; cudaEvent_t a, b; };
static std::mutex g_prof_mu;
static std::vector<EventPair> g_events;

LaunchScope::LaunchScope(int id, cudaStream_t st) : id_(id), st_(st), stop_(nullptr) {
  g_launches[id].fetch_add(1, std::memory_order_relaxed);
  if (g_profiling.load(std::memory_order_relaxed)) {
    cudaEvent_t a = nullptr, b = nullptr;
    if (cudaEventCreate(&a) == cudaSuccess && cudaEventCreate(&b) == cudaSuccess) {
      cudaEventRecord(a, st);
      stop_ = b;
      std::lock_guard<std::mutex> lk(g_prof_mu);
      g_events.push_back({id, a, b});
    }
  }
}
LaunchScope::~LaunchScope() {
  if (stop_) cudaEventRecord(stop_, st_);
}
// programmatic dependent launch between consecutive kernels of this library (bq_set_pdl).  Default OFF: measured on the headline
// step (OPT-1.3B, 232 launches, CUDA-graph replay, power-capped B200) 56.87 / 56.67 ms with it against 56.41 / 56.36 ms without
// (profiles/r02_bench_pdl_ab.json) — the step is energy-limited, the launch gaps it removes were not costing time.
static std::atomic<int> g_pdl{0};
bool pdl_enabled() { return g_pdl.load(std::memory_order_relaxed) != 0; }
}  // namespace bq

extern "C" {
int bq_kernel_count(void) { return bq::kKernCount; }
const char* bq_kernel_name(int id) { return (id >= 0 && id < bq::kKernCount) ? bq::kKernelNames[id] : ""; }
int64_t bq_launch_count(int id) {
  if (id >= 0 && id < bq::kKernCount) return bq::g_launches[id].load();
  int64_t t = 0;
  for (int i = 0; i < bq::kKernCount; ++i) t += bq::g_launches[i].load();
  return t;
}
void bq_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(bq::g_prof_mu);
  if (on) {
    for (auto& e : bq::g_events) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    bq::g_events.clear();
  }
  bq::g_profiling.store(on ? 1 : 0);
}
int bq_profile_read(int id, double* total_ms, int64_t* launches) {
  std::lock_guard<std::mutex> lk(bq::g_prof_mu);
  double tot = 0;
  int64_t n = 0;
  for (auto& e : bq::g_events) {
    if (id >= 0 && e.id != id) continue;
    if (cudaEventSynchronize(e.b) != cudaSuccess) return BQ_ERR_CUDA;
    float ms = 0;
    if (cudaEventElapsedTime(&ms, e.a, e.b) != cudaSuccess) return BQ_ERR_CUDA;
    tot += ms;
    ++n;
  }
  if (total_ms) *total_ms = tot;
  if (launches) *launches = n;
  return BQ_OK;
}
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
void bq_set_pdl(int on) { bq::g_pdl.store(on ? 1 : 0); }
int bq_get_pdl(void) { return bq::g_pdl.load(); }
int bq_abi_version(void) { return BQ_ABI_VERSION; }
const char* bq_last_cuda_error(void) { return bq::g_last_err; }
}
