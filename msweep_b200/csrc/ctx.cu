// ctx.cu — context, error state, NCCL binding (loaded lazily so that single-GPU use never needs it).
#include "common.cuh"

#include <dlfcn.h>
#include <unistd.h>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

namespace mswb {
static thread_local std::string t_last_error;
void set_last_error(const std::string &msg) { t_last_error = msg; }
std::atomic<uint64_t> g_launches{0};

// ---- device block cache (common.cuh) ------------------------------------------------------------
// Two pools per device.  Large blocks (>= 64 MB: the matrices) are parked and handed to requests of a similar size.
// Small blocks come in size classes (powers of two up to 1 MB, multiples of 1 MB beyond) and are recycled exactly: an
// optimiser run allocates ~20 K-sized vectors, a bootstrap job runs hundreds of them — cudaMalloc / cudaFree per vector
// (each cudaFree an implicit device synchronisation plus an unmap in the driver) showed up as milliseconds per replicate.
namespace {
struct CachedBlock { void *p; size_t bytes; int device; };
std::mutex g_cache_mu;
std::vector<CachedBlock> g_cache;
std::map<std::pair<int, size_t>, std::vector<void *>> g_small;      // (device, class bytes) -> parked blocks
std::map<int, size_t> g_small_held;                                  // device -> parked bytes
constexpr size_t CACHE_MIN_BYTES = (size_t)64 << 20;
constexpr size_t SMALL_CAP_BYTES = (size_t)2 << 30;                  // parked small blocks per device

size_t small_class(size_t bytes) {
  if (bytes <= ((size_t)1 << 20)) { size_t c = 512; while (c < bytes) c <<= 1; return c; }
  return (bytes + (((size_t)1 << 20) - 1)) & ~(((size_t)1 << 20) - 1);
}
} // namespace

void dev_cache_flush(int device) {
  std::vector<void *> mine;
  {
    std::lock_guard<std::mutex> lock(g_cache_mu);
    for (size_t i = 0; i < g_cache.size();)
      if (g_cache[i].device == device) { mine.push_back(g_cache[i].p); g_cache[i] = g_cache.back(); g_cache.pop_back(); } else ++i;
    for (auto &kv : g_small)
      if (kv.first.first == device) { mine.insert(mine.end(), kv.second.begin(), kv.second.end()); kv.second.clear(); }
    g_small_held[device] = 0;
  }
  for (void *p : mine) cudaFree(p);
}

void *dev_alloc(size_t bytes, size_t *capacity, int *device_out) {
  int device = 0;
  cudaGetDevice(&device);
  *device_out = device;
  if (bytes >= CACHE_MIN_BYTES) {
    std::lock_guard<std::mutex> lock(g_cache_mu);
    size_t best = g_cache.size();
    for (size_t i = 0; i < g_cache.size(); ++i) {
      const CachedBlock &b = g_cache[i];
      if (b.device == device && b.bytes >= bytes && b.bytes - bytes <= bytes / 4 && (best == g_cache.size() || b.bytes < g_cache[best].bytes)) best = i;
    }
    if (best != g_cache.size()) {
      void *p = g_cache[best].p;
      *capacity = g_cache[best].bytes;
      g_cache[best] = g_cache.back();
      g_cache.pop_back();
      return p;
    }
  } else {
    bytes = small_class(bytes);
    std::lock_guard<std::mutex> lock(g_cache_mu);
    auto it = g_small.find(std::make_pair(device, bytes));
    if (it != g_small.end() && !it->second.empty()) {
      void *p = it->second.back();
      it->second.pop_back();
      g_small_held[device] -= bytes;
      *capacity = bytes;
      return p;
    }
  }
  void *p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e == cudaErrorMemoryAllocation) {
    cudaGetLastError();
    dev_cache_flush(device);
    e = cudaMalloc(&p, bytes);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    throw Error(std::string("CUDA error: ") + cudaGetErrorString(e) + " (cudaMalloc of " + std::to_string(bytes) + " bytes)");
  }
  *capacity = bytes;
  return p;
}

// Blocks go back to the pool of the device they were allocated on — whatever device is current in the calling thread.
// Large blocks: as long as the parked total of that device stays under the cap (MSWB_CACHE_GB, default a quarter of the
// device's memory); small blocks: up to SMALL_CAP_BYTES.  mswb_ctx_trim / mswb_ctx_destroy hand everything back to the driver.
void dev_free(void *p, size_t capacity, int device) {
  if (!p) return;
  int current = 0;
  cudaGetDevice(&current);
  if (current != device) cudaSetDevice(device);
  bool parked = false;
  cudaDeviceSynchronize();   // what cudaFree would have done: nothing in flight still uses the block
  if (capacity >= CACHE_MIN_BYTES) {
    static size_t cap_bytes = 0;
    if (cap_bytes == 0) {
      if (const char *e = getenv("MSWB_CACHE_GB")) cap_bytes = (size_t)(atof(e) * 1073741824.0) + 1;
      else { size_t free_b = 0, total_b = 0; cudaMemGetInfo(&free_b, &total_b); cap_bytes = total_b / 4 + 1; }
    }
    std::lock_guard<std::mutex> lock(g_cache_mu);
    size_t held = 0;
    for (const CachedBlock &b : g_cache) if (b.device == device) held += b.bytes;
    if (held + capacity <= cap_bytes) { g_cache.push_back(CachedBlock{p, capacity, device}); parked = true; }
  } else if (capacity == small_class(capacity)) {
    std::lock_guard<std::mutex> lock(g_cache_mu);
    size_t &held = g_small_held[device];
    if (held + capacity <= SMALL_CAP_BYTES) { g_small[std::make_pair(device, capacity)].push_back(p); held += capacity; parked = true; }
  }
  if (!parked) cudaFree(p);
  if (current != device) cudaSetDevice(current);
}

// ---- large host -> device copies from pageable memory ---------------------------------------------------------------
// STAGE_THREADS host threads, each with two pinned 4 MB buffers, a stream and an event per buffer: thread t copies chunks
// t, t + T, ... into its buffers and sends them on its own stream.  Config 2's 2.7 GB CSR: 0.245 s -> 0.08 s per EC build
// (PCIe rate instead of one thread's memcpy rate).  Page-locking the buffers is a one-time cost per device (each thread pins
// its own pair, in parallel), so a cold process takes the staged path only for copies that repay it (>= 512 MB); once the
// lanes exist, every copy of 64 MB and more takes it.  MSWB_H2D_STAGED=0 turns it off.
namespace {
constexpr int STAGE_THREADS = 4;
constexpr size_t STAGE_CHUNK = (size_t)4 << 20;
constexpr size_t STAGE_COLD_MIN_BYTES = (size_t)512 << 20;
struct StageLane {
  void *buf[2] = {nullptr, nullptr};
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  bool ready = false;
  bool prepare() {                        // in the worker thread that owns the lane
    if (ready) return true;
    ready = cudaMallocHost(&buf[0], STAGE_CHUNK) == cudaSuccess && cudaMallocHost(&buf[1], STAGE_CHUNK) == cudaSuccess &&
            cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming) == cudaSuccess;
    if (!ready) cudaGetLastError();
    return ready;
  }
};
struct StageSet { StageLane lane[STAGE_THREADS]; bool warm = false; bool broken = false; std::mutex busy; };   // busy: one staged copy at a time per device
std::mutex g_stage_mu;                                   // guards the map
std::map<int, std::unique_ptr<StageSet>> g_stage;        // per device

StageSet *stage_set(int device) {
  std::lock_guard<std::mutex> map_lock(g_stage_mu);
  auto it = g_stage.find(device);
  if (it != g_stage.end()) return it->second.get();
  return (g_stage[device] = std::unique_ptr<StageSet>(new StageSet)).get();
}
} // namespace

void h2d_staged(void *dst_dev, const void *src, size_t bytes, cudaStream_t s) {
  if (bytes == 0) return;
  cudaPointerAttributes attr;
  const bool pageable = cudaPointerGetAttributes(&attr, src) != cudaSuccess || attr.type == cudaMemoryTypeUnregistered;
  cudaGetLastError();
  static const bool off = [] { const char *e = getenv("MSWB_H2D_STAGED"); return e && e[0] == '0'; }();
  int device = 0;
  cudaGetDevice(&device);
  StageSet *st = pageable && !off ? stage_set(device) : nullptr;
  static const size_t cold_min = [] { const char *e = getenv("MSWB_H2D_STAGED_MIN_MB"); return e ? (size_t)atoll(e) << 20 : STAGE_COLD_MIN_BYTES; }();
  if (!st || st->broken || (!st->warm && bytes < cold_min)) {
    MSWB_CUDA(cudaMemcpyAsync(dst_dev, src, bytes, cudaMemcpyHostToDevice, s));
    return;
  }
  std::lock_guard<std::mutex> lock(st->busy);
  MSWB_CUDA(cudaStreamSynchronize(s));          // the lanes' streams are not ordered behind s: nothing earlier may still touch dst
  const size_t n_chunks = (bytes + STAGE_CHUNK - 1) / STAGE_CHUNK;
  std::atomic<int> failed{0};
  std::atomic<size_t> next_chunk{0};            // lanes that could not be prepared leave their chunks to the others
  auto work = [&](int t) {
    StageLane &l = st->lane[t];
    if (cudaSetDevice(device) != cudaSuccess || !l.prepare()) { if (t == 0) failed = 2; return; }
    int slot = 0;
    for (size_t c = next_chunk.fetch_add(1); c < n_chunks && !failed.load(); c = next_chunk.fetch_add(1), slot ^= 1) {
      const size_t off_b = c * STAGE_CHUNK, len = std::min(STAGE_CHUNK, bytes - off_b);
      if (cudaEventSynchronize(l.ev[slot]) != cudaSuccess) { failed = 1; break; }          // the buffer's previous copy has left
      std::memcpy(l.buf[slot], static_cast<const unsigned char *>(src) + off_b, len);
      if (cudaMemcpyAsync(static_cast<unsigned char *>(dst_dev) + off_b, l.buf[slot], len, cudaMemcpyHostToDevice, l.stream) != cudaSuccess ||
          cudaEventRecord(l.ev[slot], l.stream) != cudaSuccess) { failed = 1; break; }
    }
    if (cudaStreamSynchronize(l.stream) != cudaSuccess) failed = 1;
  };
  std::thread workers[STAGE_THREADS - 1];
  for (int t = 1; t < STAGE_THREADS; ++t) workers[t - 1] = std::thread(work, t);
  work(0);
  for (auto &w : workers) w.join();
  if (failed.load() == 2) {                     // no pinned memory to be had: the plain copy, from now on
    cudaGetLastError();
    st->broken = true;
    MSWB_CUDA(cudaMemcpyAsync(dst_dev, src, bytes, cudaMemcpyHostToDevice, s));
    return;
  }
  if (failed.load()) {
    cudaGetLastError();
    throw Error("CUDA error in the staged host-to-device copy");
  }
  st->warm = true;
}
} // namespace mswb

// ---- NCCL, resolved at run time -----------------------------------------------------------------
// Only the five entry points the data path needs.  The ABI below is NCCL 2.x's (ncclUniqueId is 128
// opaque bytes passed BY VALUE; ncclDataType_t: ncclUint64 = 5, ncclFloat64 = 8; ncclSum = 0).
struct NcclUniqueId { char internal[MSWB_NCCL_ID_BYTES]; };
struct NcclApi {
  void *handle = nullptr;
  int (*GetUniqueId)(NcclUniqueId *) = nullptr;
  int (*CommInitRank)(void **, int, NcclUniqueId, int) = nullptr;
  int (*CommInitAll)(void **, int, const int *) = nullptr;
  int (*CommDestroy)(void *) = nullptr;
  int (*CommAbort)(void *) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
};

static NcclApi &nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    // Inside a torch process the already-loaded libnccl.so.2 (torch's bundled build) is reused.
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
      api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) return;
    api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.handle, "ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.handle, "ncclCommInitRank");
    api.CommInitAll = (decltype(api.CommInitAll))dlsym(api.handle, "ncclCommInitAll");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
    api.CommAbort = (decltype(api.CommAbort))dlsym(api.handle, "ncclCommAbort");
    api.AllReduce = (decltype(api.AllReduce))dlsym(api.handle, "ncclAllReduce");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
  });
  MSWB_REQUIRE(api.handle && api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce,
               "NCCL (libnccl.so.2) could not be loaded; multi-GPU contexts are unavailable");
  return api;
}

#define MSWB_NCCL(expr)                                                                              \
  do {                                                                                               \
    int _r = (expr);                                                                                 \
    if (_r != 0)                                                                                     \
      throw ::mswb::Error(std::string("NCCL error: ") +                                               \
                          (nccl().GetErrorString ? nccl().GetErrorString(_r) : "?") + " (" #expr ")"); \
  } while (0)

// ---- one-shot all-reduce over peer memory (peer.cuh) ----------------------------------------------
namespace mswb {
static __global__ void __launch_bounds__(256) peer_allreduce_kernel(double *buf, int count, PeerView pv) {
  peer_allreduce_cta<256>(buf, count, pv);
}

namespace {
// What a rank tells the others about its block (summed as 64-bit words over a zeroed table: a gather).
struct PeerRecord {
  unsigned long long pid, host, ptr, device;
  cudaIpcMemHandle_t handle;
};
static_assert(sizeof(PeerRecord) % 8 == 0, "the record travels as 64-bit words");
constexpr size_t REC_WORDS = sizeof(PeerRecord) / 8;

unsigned long long host_id() {
  char name[256] = {0};
  gethostname(name, sizeof(name) - 1);
  unsigned long long h = 1469598103934665603ull;
  for (const char *c = name; *c; ++c) { h ^= (unsigned char)*c; h *= 1099511628211ull; }
  // processes of one box share a boot id even when containers rename the host
  if (FILE *f = fopen("/proc/sys/kernel/random/boot_id", "r")) {
    char b[64] = {0};
    if (fgets(b, sizeof(b), f)) for (const char *c = b; *c; ++c) { h ^= (unsigned char)*c; h *= 1099511628211ull; }
    fclose(f);
  }
  return h;
}

bool peer_wanted() { const char *e = getenv("MSWB_PEER"); return !(e && e[0] == '0'); }
unsigned long long peer_timeout_ns() {
  double sec = 300.0;
  if (const char *e = getenv("MSWB_PEER_TIMEOUT_S")) sec = std::max(0.001, atof(e));
  return (unsigned long long)(sec * 1e9);
}

void peer_fill_view(mswb_ctx *ctx, const std::vector<void *> &blocks) {
  PeerView &v = ctx->peer;
  std::memset(&v, 0, sizeof(v));
  for (int p = 0; p < ctx->world; ++p) {
    unsigned char *b = static_cast<unsigned char *>(blocks[p]);
    v.recv[p] = reinterpret_cast<double *>(b);
    v.flags[p] = reinterpret_cast<unsigned long long *>(b + PEER_RECV_DOUBLES * 8);
  }
  unsigned long long *tail = v.flags[ctx->rank] + PEER_FLAG_WORDS;
  v.seq = tail; v.abort_word = tail + 1; v.error_word = tail + 2;
  v.timeout_ns = peer_timeout_ns();
  v.world = ctx->world; v.rank = ctx->rank;
  ctx->peer_ok = true;
}

void peer_release(mswb_ctx *ctx) {
  ctx->peer_ok = false;
  for (void *p : ctx->peer_ipc) cudaIpcCloseMemHandle(p);
  ctx->peer_ipc.clear();
  if (ctx->peer_block) cudaFree(ctx->peer_block);
  ctx->peer_block = nullptr;
  if (ctx->side_stream) cudaStreamDestroy(ctx->side_stream);
  ctx->side_stream = nullptr;
  cudaGetLastError();
}

bool peer_alloc_block(mswb_ctx *ctx) {
  if (cudaMalloc(&ctx->peer_block, PEER_BLOCK_BYTES) != cudaSuccess) { cudaGetLastError(); ctx->peer_block = nullptr; return false; }
  if (cudaMemset(ctx->peer_block, 0, PEER_BLOCK_BYTES) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return true;
}

// One process per GPU (or one thread per GPU with its own mswb_ctx_create): the ranks gather each other's records over
// the NCCL communicator they already share, map the blocks (CUDA IPC across processes, peer access inside one), and
// agree — with one more sum — whether EVERY rank succeeded; otherwise all of them keep NCCL for the data path.
void peer_setup_exchange(mswb_ctx *ctx) {
  if (!peer_wanted() || ctx->world > PEER_MAX_WORLD) return;
  const int W = ctx->world;
  unsigned long long failed = peer_alloc_block(ctx) ? 0ull : 1ull;
  PeerRecord mine;
  std::memset(&mine, 0, sizeof(mine));
  mine.pid = (unsigned long long)getpid(); mine.host = host_id();
  mine.ptr = (unsigned long long)(uintptr_t)ctx->peer_block; mine.device = (unsigned long long)ctx->device;
  if (!failed && cudaIpcGetMemHandle(&mine.handle, ctx->peer_block) != cudaSuccess) { cudaGetLastError(); failed = 1ull; }
  const size_t words = (size_t)W * REC_WORDS + 1;
  std::vector<unsigned long long> table(words, 0ull);
  std::memcpy(&table[(size_t)ctx->rank * REC_WORDS], &mine, sizeof(mine));
  table[words - 1] = failed;
  DevBuf<unsigned long long> dev;
  dev.alloc(words);
  h2d(dev.p, table.data(), words, ctx->stream);
  ctx->allreduce_sum_u64(dev.p, words);
  d2h(table.data(), dev.p, words, ctx->stream);
  MSWB_CUDA(cudaStreamSynchronize(ctx->stream));
  bool ok = table[words - 1] == 0ull;
  std::vector<void *> blocks(W, nullptr);
  if (ok) {
    for (int p = 0; p < W && ok; ++p) {
      PeerRecord r;
      std::memcpy(&r, &table[(size_t)p * REC_WORDS], sizeof(r));
      if (p == ctx->rank) { blocks[p] = ctx->peer_block; continue; }
      if (r.host != mine.host) { ok = false; break; }
      if (r.pid == mine.pid) {
        // another context of this process: plain peer access
        int can = 0;
        if ((int)r.device == ctx->device) { blocks[p] = (void *)(uintptr_t)r.ptr; continue; }
        if (cudaDeviceCanAccessPeer(&can, ctx->device, (int)r.device) != cudaSuccess || !can) { cudaGetLastError(); ok = false; break; }
        const cudaError_t e = cudaDeviceEnablePeerAccess((int)r.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); ok = false; break; }
        cudaGetLastError();
        blocks[p] = (void *)(uintptr_t)r.ptr;
      } else {
        void *mapped = nullptr;
        if (cudaIpcOpenMemHandle(&mapped, r.handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
        ctx->peer_ipc.push_back(mapped);
        blocks[p] = mapped;
      }
    }
  }
  // second round: did every rank map every block?
  unsigned long long bad = ok ? 0ull : 1ull;
  h2d(dev.p, &bad, 1, ctx->stream);
  ctx->allreduce_sum_u64(dev.p, 1);
  d2h(&bad, dev.p, 1, ctx->stream);
  MSWB_CUDA(cudaStreamSynchronize(ctx->stream));
  if (bad != 0ull) { peer_release(ctx); return; }
  peer_fill_view(ctx, blocks);
}

// One host thread per GPU inside one process (mswb_ctx_create_group): no exchange needed.
void peer_setup_group(int n, mswb_ctx **ctxs) {
  if (!peer_wanted() || n < 2 || n > PEER_MAX_WORLD) return;
  bool ok = true;
  for (int i = 0; i < n && ok; ++i) {
    if (cudaSetDevice(ctxs[i]->device) != cudaSuccess) { ok = false; break; }
    ok = peer_alloc_block(ctxs[i]);
    for (int j = 0; j < n && ok; ++j) {
      if (j == i || ctxs[j]->device == ctxs[i]->device) continue;
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, ctxs[i]->device, ctxs[j]->device) != cudaSuccess || !can) { ok = false; break; }
      const cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[j]->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = false;
      cudaGetLastError();
    }
  }
  if (!ok) {
    cudaGetLastError();
    for (int i = 0; i < n; ++i) { cudaSetDevice(ctxs[i]->device); peer_release(ctxs[i]); }
    return;
  }
  std::vector<void *> blocks(n);
  for (int i = 0; i < n; ++i) blocks[i] = ctxs[i]->peer_block;
  for (int i = 0; i < n; ++i) peer_fill_view(ctxs[i], blocks);
}
} // namespace
} // namespace mswb

void mswb_ctx::peer_check() {
  if (!peer_ok) return;
  unsigned long long err = 0ull;
  MSWB_CUDA(cudaMemcpy(&err, peer.error_word, sizeof(err), cudaMemcpyDeviceToHost));
  MSWB_REQUIRE(err == 0ull, "a peer rank did not arrive at the all-reduce (it failed, was aborted, or the wait timed out: MSWB_PEER_TIMEOUT_S)");
}

void mswb_ctx::allreduce_sum(double *buf_dev, size_t count) {
  if (world == 1 || count == 0) return;
  if (peer_ok && count <= (size_t)mswb::PEER_SLOT_DOUBLES) {
    mswb::peer_allreduce_kernel<<<1, 256, 0, stream>>>(buf_dev, (int)count, peer);
    MSWB_LAUNCHED();
    return;
  }
  void *comm = nccl_comm.load();
  MSWB_REQUIRE(comm, "the communicator of this context has been aborted (a peer rank failed)");
  MSWB_NCCL(nccl().AllReduce(buf_dev, buf_dev, count, /*ncclFloat64*/ 8, /*ncclSum*/ 0, comm, stream));
}
void mswb_ctx::allreduce_sum_u64(unsigned long long *buf_dev, size_t count) {
  if (world == 1 || count == 0) return;
  void *comm = nccl_comm.load();
  MSWB_REQUIRE(comm, "the communicator of this context has been aborted (a peer rank failed)");
  MSWB_NCCL(nccl().AllReduce(buf_dev, buf_dev, count, /*ncclUint64*/ 5, /*ncclSum*/ 0, comm, stream));
}

extern "C" {

const char *mswb_last_error(void) { return mswb::t_last_error.c_str(); }
const char *mswb_version(void) { return "msweep-b200 0.1.0 (sm_100a)"; }
uint64_t mswb_launch_count(void) { return mswb::g_launches.load(); }

int mswb_nccl_unique_id(void *out_id) {
  return mswb::guarded([&] {
    MSWB_REQUIRE(out_id, "out_id is NULL");
    NcclUniqueId id;
    MSWB_NCCL(nccl().GetUniqueId(&id));
    std::memcpy(out_id, &id, sizeof(id));
  });
}

int mswb_ctx_create(int device, int rank, int world_size, const void *nccl_id, void *cuda_stream, mswb_ctx **out) {
  return mswb::guarded([&] {
    MSWB_REQUIRE(out, "out is NULL");
    MSWB_REQUIRE(world_size >= 1 && rank >= 0 && rank < world_size, "bad rank / world_size");
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    MSWB_REQUIRE(e == cudaSuccess && n_dev > 0,
                 "no CUDA device available: msweep_b200 has no CPU fallback (this backend needs an sm_100a GPU)");
    MSWB_REQUIRE(device >= 0 && device < n_dev, "device index out of range");
    MSWB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    MSWB_CUDA(cudaGetDeviceProperties(&prop, device));
    MSWB_REQUIRE(prop.major == 10, std::string("device is sm_") + std::to_string(prop.major) + std::to_string(prop.minor) +
                                       "; this library carries sm_100a code only");
    std::unique_ptr<mswb_ctx> ctx(new mswb_ctx);
    ctx->device = device; ctx->rank = rank; ctx->world = world_size;
    ctx->n_sms = prop.multiProcessorCount;
    if (cuda_stream) { ctx->stream = (cudaStream_t)cuda_stream; ctx->own_stream = false; }
    else { MSWB_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)); ctx->own_stream = true; }
    if (world_size > 1) {
      MSWB_REQUIRE(nccl_id, "nccl_id is required when world_size > 1");
      NcclUniqueId id;
      std::memcpy(&id, nccl_id, sizeof(id));
      void *comm = nullptr;
      MSWB_NCCL(nccl().CommInitRank(&comm, world_size, id, rank));
      ctx->nccl_comm = comm;
      mswb::peer_setup_exchange(ctx.get());
    }
    *out = ctx.release();
  });
}

int mswb_ctx_create_group(int n, const int *devices, mswb_ctx **out) {
  return mswb::guarded([&] {
    MSWB_REQUIRE(n >= 1 && devices && out, "bad arguments");
    std::vector<void *> comms(n, nullptr);
    if (n > 1) {
      MSWB_REQUIRE(nccl().CommInitAll, "ncclCommInitAll is not available in the loaded NCCL");
      MSWB_NCCL(nccl().CommInitAll(comms.data(), n, devices));
    }
    for (int i = 0; i < n; ++i) {
      out[i] = nullptr;
      if (mswb_ctx_create(devices[i], 0, 1, nullptr, nullptr, &out[i])) throw mswb::Error(mswb_last_error());
      out[i]->rank = i;
      out[i]->world = n;
      out[i]->nccl_comm = comms[i];
    }
    mswb::peer_setup_group(n, out);
  });
}

int mswb_ctx_abort(mswb_ctx *ctx) {
  return mswb::guarded([&] {
    MSWB_REQUIRE(ctx, "ctx is NULL");
    if (ctx->peer_ok && ctx->side_stream) {
      // a kernel of this rank may be waiting for a peer that will never push: end the wait
      int cur = 0;
      cudaGetDevice(&cur);
      cudaSetDevice(ctx->device);
      cudaMemsetAsync(ctx->peer.abort_word, 0xff, sizeof(unsigned long long), ctx->side_stream);
      cudaStreamSynchronize(ctx->side_stream);
      cudaSetDevice(cur);
      cudaGetLastError();
    }
    void *comm = ctx->nccl_comm.exchange(nullptr);
    if (!comm) return;
    MSWB_REQUIRE(nccl().CommAbort, "ncclCommAbort is not available in the loaded NCCL");
    MSWB_NCCL(nccl().CommAbort(comm));     // pending collectives of this rank end with an error instead of waiting for ever
  });
}

int mswb_ctx_peer_active(const mswb_ctx *ctx) { return ctx && ctx->peer_ok ? 1 : 0; }

void mswb_ctx_destroy(mswb_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (void *comm = ctx->nccl_comm.exchange(nullptr)) nccl().CommDestroy(comm);
  mswb::peer_release(ctx);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  mswb::dev_cache_flush(ctx->device);
  delete ctx;
}

int mswb_ctx_trim(mswb_ctx *ctx) {
  return mswb::guarded([&] {
    MSWB_REQUIRE(ctx, "ctx is NULL");
    MSWB_CUDA(cudaSetDevice(ctx->device));
    MSWB_CUDA(cudaStreamSynchronize(ctx->stream));
    mswb::dev_cache_flush(ctx->device);
  });
}

int mswb_device_warmup(int device) {
  return mswb::guarded([&] {
    int n_dev = 0;
    MSWB_REQUIRE(cudaGetDeviceCount(&n_dev) == cudaSuccess && n_dev > 0, "no CUDA device available: msweep_b200 has no CPU fallback");
    MSWB_REQUIRE(device >= 0 && device < n_dev, "device index out of range");
    MSWB_CUDA(cudaSetDevice(device));
    MSWB_CUDA(cudaFree(nullptr));
  });
}

int mswb_ctx_sync(mswb_ctx *ctx) {
  return mswb::guarded([&] {
    MSWB_REQUIRE(ctx, "ctx is NULL");
    MSWB_CUDA(cudaSetDevice(ctx->device));
    MSWB_CUDA(cudaStreamSynchronize(ctx->stream));
  });
}

int mswb_shard_range(const mswb_ctx *ctx, uint64_t n_ecs, uint64_t *begin, uint64_t *end) {
  return mswb::guarded([&] {
    MSWB_REQUIRE(ctx && begin && end, "NULL argument");
    // contiguous, balanced to within one class: rank r owns [floor(N r / W), floor(N (r+1) / W))
    *begin = (uint64_t)((unsigned __int128)n_ecs * ctx->rank / ctx->world);
    *end = (uint64_t)((unsigned __int128)n_ecs * (ctx->rank + 1) / ctx->world);
  });
}

} // extern "C"
