// bq_internal.h — shared host-side declarations of libbq_b200.so (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>

#include "../../include/bq.h"

namespace bq {
void set_last_cuda_error(const char* what, const char* file, int line);
int num_sms();
int quantize_impl(const bq_format* fmt, const bq_tensor3* t, const float* x, void* y, int y_dtype, int transpose_out,
                  void* ws, size_t ws_bytes, cudaStream_t st, const float* x2 = nullptr);
size_t quantize_ws_bytes(const bq_format* fmt, const bq_tensor3* t);
int gemm_bf16_tn_impl(const void* A, const void* B, float* C, const float* bias, int64_t batch, int64_t M, int64_t N,
                      int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int64_t sa, int64_t sb, int64_t sc,
                      cudaStream_t st);
struct FmtParams;
int make_params(const bq_format* f, FmtParams* p);
// Per-device slot for lazily initialised launch state (function attributes and occupancy are per-device properties: a process
// that drives more than one GPU must not reuse device 0's answers).
int current_device();
template <typename T> struct PerDevice {
  T v[64] = {};
  T& get() { return v[current_device()]; }
};
enum KernelId { kKernQuantRows = 0, kKernBlockLogFixup, kKernQuantTile, kKernGenericMax, kKernGenericMin, kKernGenericQuant,
                kKernGemm, kKernAttention, kKernSplit3, kKernGemmEpi, kKernGemmSplit, kKernLnQuant, kKernQuantStream, kKernSiluMulQuant, kKernTokenCe, kKernTokenCeMean, kKernPeerBarrier, kKernRopeQuant, kKernPeerPush, kKernGemmXformA, kKernGemmXformB, kKernPackWeight, kKernSoftmaxQuant, kKernRopeSplit, kKernSplit3T, kKernCount };
// RAII launch bracket: counts the launch; records start/stop events on `st` when profiling is enabled.
struct LaunchScope {
  LaunchScope(int id, cudaStream_t st);
  ~LaunchScope();
  int id_;
  cudaStream_t st_;
  cudaEvent_t stop_;
};
// Kernel launch through cudaLaunchKernelEx with the optional attributes our kernels use: thread-block cluster width and
// programmatic dependent launch (see ptx::griddep_wait; on when pdl_enabled()).  Also valid during stream capture (the dependency
// becomes a programmatic edge of the CUDA graph).
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_ex(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, int cluster, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3(block, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  int n = 0;
  if (cluster > 1) {
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = cluster; at[n].val.clusterDim.y = 1; at[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = at;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}
}  // namespace bq

#define BQ_CUDA_CHECK(expr)                                                  \
  do {                                                                       \
    cudaError_t _e = (expr);                                                 \
    if (_e != cudaSuccess) {                                                 \
      bq::set_last_cuda_error(cudaGetErrorString(_e), __FILE__, __LINE__);   \
      return BQ_ERR_CUDA;                                                    \
    }                                                                        \
  } while (0)
