// peer.cu — peer-memory plumbing for the column-parallel Linear (the reference runs on one device and has
// no counterpart): legacy CUDA IPC export / import of caller-owned buffers and a flag barrier over peer-mapped memory.
// The data movement itself is not here: the GEMM epilogue (gemm_sm100.cu, EpiArgs::rep) stores every output tile straight
// into the peers' buffers over NVLink, so the all-gather overlaps the multiply tile by tile; this file only orders it.
#include "bq_internal.h"

#include <cuda.h>

namespace bq {

typedef CUresult (*PFN_getAddressRange)(CUdeviceptr*, size_t*, CUdeviceptr);
static PFN_getAddressRange g_get_range = nullptr;

static int load_get_range() {
  if (g_get_range) return BQ_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  BQ_CUDA_CHECK(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qres));
  if (qres != cudaDriverEntryPointSuccess || !fn) {
    set_last_cuda_error("cuMemGetAddressRange entry point not available", __FILE__, __LINE__);
    return BQ_ERR_CUDA;
  }
  g_get_range = (PFN_getAddressRange)fn;
  return BQ_OK;
}

__device__ __forceinline__ void red_release_sys_add(uint32_t* p, uint32_t v) {
  asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

struct PeerSignals { uint32_t* blk[8]; };

// One CTA, one thread per peer.  Thread p: (1) system-scope release-increment of word [rank] in peer p's block — every
// write this GPU issued before the barrier (kernel boundary + the release) is visible to p once p observes the count;
// (2) acquire-wait for word [p] of the own block to reach `epoch`.  Counters only grow, so there is no reset race; the
// comparison is on the signed difference (wrap-safe).
// epoch == 0: the count lives in word [BQ_PEER_FLAG_EPOCH] of the own block and is advanced by the kernel — the launch carries no
// per-call host state, so it can be captured in a CUDA graph and replayed (a host-side epoch would be frozen at capture time and
// every replayed barrier would pass at once).
__global__ void peer_barrier_kernel(PeerSignals s, int rank, int world, uint32_t epoch, uint64_t timeout_ns, uint32_t* host_flag) {
  const int p = threadIdx.x;
  const bool dev_epoch = epoch == 0u;
  if (dev_epoch) {
    __shared__ uint32_t e;
    if (p == 0) e = s.blk[rank][BQ_PEER_FLAG_EPOCH] + 1u;
    __syncthreads();
    epoch = e;
    __syncthreads();
    if (p == 0) s.blk[rank][BQ_PEER_FLAG_EPOCH] = epoch;          // only this rank's barrier kernels (stream-ordered) touch the word
  }
  if (p < world && p != rank) {
    __threadfence_system();
    red_release_sys_add(s.blk[p] + rank, 1u);
    const uint32_t* mine = s.blk[rank] + p;
    const uint64_t t0 = globaltimer_ns();
    while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
      __nanosleep(64);
      if (globaltimer_ns() - t0 > timeout_ns) {
        s.blk[rank][BQ_PEER_FLAG_TIMEOUT] = 1u;
        if (host_flag) {                                        // sticky, in mapped pinned host memory: the host sees it without a sync
          *reinterpret_cast<volatile uint32_t*>(host_flag) = 1u;
          __threadfence_system();
        }
        break;
      }
    }
  }
}

// Copy a strided [rows x row_bytes] slab of local memory to the same position of up to 7 peer-mapped buffers (the all-gather of a
// slab that was NOT produced by a GEMM epilogue: the attention output of this rank's heads).  Every warp instruction stores
// 32 x 16 contiguous bytes; each source element is read once and written n_dst times.  Grid-stride over 16-byte packets.
struct PushArgs {
  const uint4* src;
  uint4* dst[BQ_MAX_REPLICAS];
  int n_dst;
  int64_t rows, packets_per_row, src_stride, dst_stride;      // strides in 16-byte packets
};
__global__ void __launch_bounds__(256) peer_push_kernel(PushArgs a) {
  const int64_t total = a.rows * a.packets_per_row;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / a.packets_per_row, c = i - r * a.packets_per_row;
    const uint4 v = a.src[r * a.src_stride + c];
#pragma unroll 1
    for (int p = 0; p < a.n_dst; ++p) a.dst[p][r * a.dst_stride + c] = v;
  }
}

}  // namespace bq

extern "C" {

int bq_peer_push(const void* src, void* const* dst, int32_t n_dst, int64_t rows, int64_t row_bytes, int64_t src_stride_bytes,
                 int64_t dst_stride_bytes, void* stream) {
  if (n_dst < 0 || n_dst > BQ_MAX_REPLICAS || rows < 0 || row_bytes < 0) return BQ_ERR_BAD_ARG;
  if (n_dst == 0 || rows == 0 || row_bytes == 0) return BQ_OK;
  if (!src || !dst || (row_bytes % 16) || (src_stride_bytes % 16) || (dst_stride_bytes % 16) || ((uintptr_t)src % 16) ||
      src_stride_bytes < row_bytes || dst_stride_bytes < row_bytes)
    return BQ_ERR_BAD_ARG;
  bq::PushArgs a;
  memset(&a, 0, sizeof(a));
  a.src = (const uint4*)src;
  for (int i = 0; i < n_dst; ++i) {
    if (!dst[i] || ((uintptr_t)dst[i] % 16)) return BQ_ERR_BAD_ARG;
    a.dst[i] = (uint4*)dst[i];
  }
  a.n_dst = n_dst; a.rows = rows; a.packets_per_row = row_bytes / 16;
  a.src_stride = src_stride_bytes / 16; a.dst_stride = dst_stride_bytes / 16;
  const int64_t total = rows * a.packets_per_row;
  const int grid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)bq::num_sms() * 8);
  cudaStream_t st = (cudaStream_t)stream;
  {
    bq::LaunchScope ls(bq::kKernPeerPush, st);
    bq::peer_push_kernel<<<grid, 256, 0, st>>>(a);
  }
  BQ_CUDA_CHECK(cudaGetLastError());
  return BQ_OK;
}

int bq_ipc_export(const void* dev_ptr, bq_ipc_handle* out) {
  if (!dev_ptr || !out) return BQ_ERR_BAD_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  int rc = bq::load_get_range();
  if (rc) return rc;
  CUdeviceptr base = 0;
  size_t size = 0;
  CUresult r = bq::g_get_range(&base, &size, (CUdeviceptr)dev_ptr);
  if (r != CUDA_SUCCESS) {
    char msg[96];
    snprintf(msg, sizeof(msg), "cuMemGetAddressRange failed with CUresult %d", (int)r);
    bq::set_last_cuda_error(msg, __FILE__, __LINE__);
    return BQ_ERR_CUDA;
  }
  cudaIpcMemHandle_t h;
  BQ_CUDA_CHECK(cudaIpcGetMemHandle(&h, (void*)base));
  memcpy(out->reserved, &h, 64);
  out->offset = (int64_t)((CUdeviceptr)dev_ptr - base);
  out->size = (int64_t)size;
  return BQ_OK;
}

int bq_ipc_import(const bq_ipc_handle* h, void** base, void** ptr) {
  if (!h || !base || !ptr || h->offset < 0 || h->offset > h->size) return BQ_ERR_BAD_ARG;
  cudaIpcMemHandle_t ch;
  memcpy(&ch, h->reserved, 64);
  void* b = nullptr;
  BQ_CUDA_CHECK(cudaIpcOpenMemHandle(&b, ch, cudaIpcMemLazyEnablePeerAccess));
  *base = b;
  *ptr = (char*)b + h->offset;
  return BQ_OK;
}

int bq_ipc_release(void* base) {
  if (!base) return BQ_ERR_BAD_ARG;
  BQ_CUDA_CHECK(cudaIpcCloseMemHandle(base));
  return BQ_OK;
}

int bq_peer_barrier(void* const* signals, int32_t rank, int32_t world, uint32_t epoch, int32_t timeout_ms, void* stream) {
  return bq_peer_barrier_ex(signals, rank, world, epoch, timeout_ms, nullptr, stream);
}

int bq_peer_barrier_ex(void* const* signals, int32_t rank, int32_t world, uint32_t epoch, int32_t timeout_ms, uint32_t* host_error_flag,
                       void* stream) {
  if (!signals || world < 1 || world > 8 || rank < 0 || rank >= world || timeout_ms <= 0) return BQ_ERR_BAD_ARG;
  bq::PeerSignals s;
  for (int i = 0; i < 8; ++i) {
    s.blk[i] = i < world ? (uint32_t*)signals[i] : nullptr;
    if (i < world && (!s.blk[i] || ((uintptr_t)s.blk[i] & 3))) return BQ_ERR_BAD_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  {
    bq::LaunchScope ls(bq::kKernPeerBarrier, st);
    bq::peer_barrier_kernel<<<1, 32, 0, st>>>(s, rank, world, epoch, (uint64_t)timeout_ms * 1000000ull, host_error_flag);
  }
  BQ_CUDA_CHECK(cudaGetLastError());
  return BQ_OK;
}

}  // extern "C"
