// common.cuh — shared host/device helpers of libmsweep_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <map>
#include <memory>
#include <mutex>
#include <algorithm>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/msweep_b200.h"
#include "peer.cuh"

namespace mswb {

// ---- errors -----------------------------------------------------------------------------------
void set_last_error(const std::string &msg);
struct Error : std::runtime_error { using std::runtime_error::runtime_error; };

#define MSWB_CUDA(expr)                                                                              \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess && cudaGetLastError() != (cudaError_t)-1) /* (clears the recorded error) */ \
      throw ::mswb::Error(std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " + __FILE__ +  \
                          ":" + std::to_string(__LINE__) + " (" #expr ")");                          \
  } while (0)

#define MSWB_REQUIRE(cond, msg)                                                                      \
  do { if (!(cond)) throw ::mswb::Error(msg); } while (0)

template <typename F> int guarded(F &&f) {
  try { f(); return 0; }
  catch (const std::exception &e) { set_last_error(e.what()); return 1; }
  catch (...) { set_last_error("unknown error"); return 1; }
}

extern std::atomic<uint64_t> g_launches;
// Counts a kernel launch of OURS and checks the launch error.
#define MSWB_LAUNCHED()                                                                              \
  do { ::mswb::g_launches.fetch_add(1, std::memory_order_relaxed); MSWB_CUDA(cudaGetLastError()); } while (0)

// ---- device buffers ---------------------------------------------------------------------------
// cudaMalloc / cudaFree of the multi-GB matrices cost tens to hundreds of milliseconds (page tables, an implicit device
// synchronisation): blocks of 64 MB and more go back to a small per-device cache instead and are handed out again to
// requests of a similar size.  The cache is emptied when an allocation fails and when a context is destroyed (ctx.cu).
void *dev_alloc(size_t bytes, size_t *capacity, int *device);
void dev_free(void *p, size_t capacity, int device);   // device: the one the block was allocated on
void dev_cache_flush(int device);

template <typename T> struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  size_t cap = 0;   // bytes of the underlying block
  int dev = 0;      // device the block lives on
  DevBuf() = default;
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n), cap(o.cap), dev(o.dev) { o.p = nullptr; o.n = 0; o.cap = 0; }
  DevBuf &operator=(DevBuf &&o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; cap = o.cap; dev = o.dev; o.p = nullptr; o.n = 0; o.cap = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void release() { if (p) dev_free(p, cap, dev); p = nullptr; n = 0; cap = 0; }
  void alloc(size_t count) {
    release();
    if (count == 0) count = 1;
    p = static_cast<T *>(dev_alloc(count * sizeof(T), &cap, &dev));
    n = count;
  }
  void ensure(size_t count) { if (count > n) alloc(count); }
  size_t bytes() const { return n * sizeof(T); }
};

template <typename T> struct PinnedBuf {
  T *p = nullptr;
  size_t n = 0;
  PinnedBuf() = default;
  PinnedBuf(const PinnedBuf &) = delete;
  PinnedBuf &operator=(const PinnedBuf &) = delete;
  ~PinnedBuf() { if (p) cudaFreeHost(p); }
  void alloc(size_t count) {
    if (p) cudaFreeHost(p);
    p = nullptr;
    if (count == 0) count = 1;
    MSWB_CUDA(cudaMallocHost((void**)&p, count * sizeof(T)));
    n = count;
  }
  void ensure(size_t count) { if (count > n) alloc(count); }
};

// Host -> device.  Small copies and copies from page-locked memory are plain async copies on the stream.  A LARGE copy from
// pageable memory (the caller's std::vector of a multi-GB pseudoalignment) is what cudaMemcpy does slowly — one host thread
// staging through the driver's bounce buffer, 6-10 GB/s: h2d_staged (ctx.cu) cuts it into chunks that several host threads
// copy into their own pinned buffers and send on their own streams, and returns with the data on the device.
void h2d_staged(void *dst_dev, const void *src, size_t bytes, cudaStream_t s);
constexpr size_t H2D_STAGED_MIN_BYTES = (size_t)64 << 20;
template <typename T> void h2d(T *dst_dev, const T *src, size_t count, cudaStream_t s) {
  if (!count) return;
  if (count * sizeof(T) >= H2D_STAGED_MIN_BYTES) { h2d_staged(dst_dev, src, count * sizeof(T), s); return; }
  MSWB_CUDA(cudaMemcpyAsync(dst_dev, src, count * sizeof(T), cudaMemcpyHostToDevice, s));
}
template <typename T> void d2h(T *dst, const T *src_dev, size_t count, cudaStream_t s) {
  if (count) MSWB_CUDA(cudaMemcpyAsync(dst, src_dev, count * sizeof(T), cudaMemcpyDeviceToHost, s));
}

// Kernels that take more than 48 KB of dynamic shared memory need the function attribute raised first.  It is raised, never
// lowered (a smaller request of another problem must not take it away from a cached larger one), per (kernel, device).
template <class Kern> void ensure_dyn_smem(Kern kern, int device, size_t smem) {
  if (smem <= 48 * 1024) return;
  static std::map<std::pair<const void *, int>, size_t> raised;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  size_t &have = raised[std::make_pair((const void *)kern, device)];
  if (smem > have) {
    MSWB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    have = smem;
  }
}

inline uint64_t ceil_div(uint64_t a, uint64_t b) { return (a + b - 1) / b; }
// The linear-domain likelihood, the counts and the row maxima are padded with zero rows to a multiple of ROW_PAD
// classes (>= the rows of one CTA batch of every tile shape): the EM sweep carries no row-bounds predicates.
constexpr int ROW_PAD = 256;
inline uint64_t round_up(uint64_t a, uint64_t b) { return ceil_div(a, b) * b; }

// ---- device math ------------------------------------------------------------------------------
#ifdef __CUDACC__
// The authors' digamma (recurrence to x >= 7, then a series in 1/(x - 1/2)); the same function the
// reference carries in src/Sample.cpp:87-97 and rcgpar uses in its optimiser.
__host__ __device__ inline double digamma_series(double x) {
  double acc = 0.0;
  while (x < 7.0) { acc -= 1.0 / x; x += 1.0; }
  x -= 0.5;
  const double r = 1.0 / x, r2 = r * r, r4 = r2 * r2;
  acc += log(x) + (1.0 / 24.0) * r2 - (7.0 / 960.0) * r4 + (31.0 / 8064.0) * r4 * r2 - (127.0 / 30720.0) * r4 * r4;
  return acc;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum in a fixed order (deterministic); every thread gets the result.
// scratch: at least 32 doubles of shared memory, reusable after the call returns.
template <int NT> __device__ __forceinline__ double block_sum(double v, double *scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int w = 0; w < NT / 32; ++w) s += scratch[w];
  return s;
}
template <int NT> __device__ __forceinline__ double block_max(double v, double *scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double s = scratch[0];
#pragma unroll
  for (int w = 1; w < NT / 32; ++w) s = fmax(s, scratch[w]);
  return s;
}

// 16-byte streaming loads/stores: the likelihood / gamma / step arrays are touched once per sweep
// and are far larger than L2, so they should not displace the small per-group vectors in L1.
__device__ __forceinline__ double2 ld_stream(const double2 *p) { return __ldcs(p); }
__device__ __forceinline__ float4 ld_stream(const float4 *p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(double2 *p, double2 v) { __stcs(p, v); }
#endif

} // namespace mswb

// ---- handle definitions (shared by the translation units) ---------------------------------------
struct NcclApi;   // dynamically loaded, see ctx.cu

struct mswb_ctx {
  int device = 0, rank = 0, world = 1;
  int n_sms = 148;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::atomic<void *> nccl_comm{nullptr};   // (mswb_ctx_abort may take it away from another thread)
  mswb::DevBuf<double> comm_buf;       // staging for the per-pass all-reduce
  // One-shot all-reduce over NVLink peer memory (peer.cuh): set up at context creation when every rank of the world can
  // map every other rank's block (one process per GPU: CUDA IPC; one thread per GPU: peer access); else NCCL carries it.
  bool peer_ok = false;
  mswb::PeerView peer{};
  void *peer_block = nullptr;               // this rank's receive area + flags (cudaMalloc)
  std::vector<void *> peer_ipc;             // blocks of other processes opened through CUDA IPC
  cudaStream_t side_stream = nullptr;       // for the abort word (must not queue behind a waiting kernel)
  void allreduce_sum(double *buf_dev, size_t count);           // no-op when world == 1; peer memory or NCCL
  void allreduce_sum_u64(unsigned long long *buf_dev, size_t count);   // NCCL
  void peer_check();                        // throws when a peer exchange gave up (call after a stream synchronisation)
};
