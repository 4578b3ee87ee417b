// vi.cu — variational optimiser: control kernels, tile dispatch and the host-side session.
// Replaces rcgpar::rcg_optl_omp / rcg_optl_torch / em_torch / mixture_components as called from the
// reference at src/mSWEEP.cpp:192-203, 419-423, 507-516.
#include "handles.cuh"
#include "vi_kernels.cuh"

#include <cmath>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <mutex>
#include <tuple>

using namespace mswb;

// =====================================================================================================
// control kernels (one CTA; K-sized work)
// =====================================================================================================
namespace mswb {

constexpr int CTL_NT = 256;

struct ViArrays {
  double *alpha0, *N_k, *dg, *w;   // [K]   dg = digamma(N_k) (EM) or digamma(N_k) - 1 (RCG)
  double *dg_prev;                  // [K]   EM: the digamma vector the LAST pass used (posteriors on demand)
  double *red;                      // [K + 2] reduced sums of the last sweep
  double *trace_bound, *trace_gnorm;
  unsigned char *trace_reset;
  unsigned long long trace_cap;
};

__device__ __forceinline__ void trace_push(const ViArrays &a, ViCtl *ctl, double gnorm, int reset) {
  if (ctl->iter < a.trace_cap) {
    a.trace_bound[ctl->iter] = ctl->bound;
    a.trace_gnorm[ctl->iter] = gnorm;
    a.trace_reset[ctl->iter] = (unsigned char)reset;
  }
}

// N_k = alpha0 + total/K (what gamma = log(1/K) gives), first digamma vector, control block reset.
__global__ void vi_init_kernel(ViArrays a, ViCtl *ctl, int K, int algo, double tol, unsigned long long max_iters,
                               double bound_const, double sum_counts) {
  __shared__ double scratch[32];
  double mx = -INFINITY;
  for (int k = threadIdx.x; k < K; k += CTL_NT) {
    const double nk = a.alpha0[k] + sum_counts / (double)K;
    a.N_k[k] = nk;
    const double dg = digamma_series(nk);
    a.dg[k] = algo == MSWB_ALGO_RCG ? dg - 1.0 : dg;
    mx = fmax(mx, dg);
  }
  mx = block_max<CTL_NT>(mx, scratch);
  for (int k = threadIdx.x; k < K; k += CTL_NT) a.w[k] = exp(a.dg[k] - mx);   // EM only
  if (threadIdx.x == 0) {
    ctl->bound = algo == MSWB_ALGO_RCG ? -100000.0 : 0.0;
    ctl->oldbound = ctl->bound;
    ctl->oldnorm = 1.0; ctl->newnorm = 0.0; ctl->beta = 0.0;
    ctl->bound_const = bound_const; ctl->tol = tol; ctl->sum_counts = sum_counts; ctl->dg_max = mx;
    ctl->iter = 0; ctl->max_iters = max_iters; ctl->resets = 0;
    ctl->use_old = 0; ctl->didreset = 0; ctl->converged = 0; ctl->fault = 0;
    ctl->done = max_iters == 0 ? 1 : 0;
  }
}

// EM: N_k, bound, convergence test, then the next digamma / weight vector.
// red[k] = sum_j P(j,k) c_j / S_j (without w_k), red[K] = sum_j c_j (log S_j + M_j).
__global__ void em_ctl_kernel(ViArrays a, ViCtl *ctl, int K, int linear) {
  if (ctl->done) return;
  __shared__ double scratch[32];
  double lg = 0.0, dga = 0.0;
  for (int k = threadIdx.x; k < K; k += CTL_NT) {
    const double A = linear ? a.w[k] * (a.red[k] + a.red[K + 1]) : a.red[k];   // red[K+1]: the share every group gets (sparse pass), else 0
    const double nk = a.alpha0[k] + A;
    a.N_k[k] = nk;
    lg += lgamma(nk);
    dga += a.dg[k] * A;
  }
  lg = block_sum<CTL_NT>(lg, scratch);
  dga = block_sum<CTL_NT>(dga, scratch);
  double mx = -INFINITY;
  for (int k = threadIdx.x; k < K; k += CTL_NT) {
    const double dg = digamma_series(a.N_k[k]);
    a.dg_prev[k] = a.dg[k];
    a.dg[k] = dg;
    mx = fmax(mx, dg);
  }
  mx = block_max<CTL_NT>(mx, scratch);
  for (int k = threadIdx.x; k < K; k += CTL_NT) a.w[k] = exp(a.dg[k] - mx);
  if (threadIdx.x == 0) {
    // linear: log-normaliser of class j is log S_j + M_j + dg_max, and sum_k q (logl - gamma) = lse_j - sum_k q dg_k
    const double data = linear ? a.red[K] + ctl->sum_counts * ctl->dg_max - dga : a.red[K];
    const double bound = data + lg + ctl->bound_const;
    ctl->oldbound = ctl->bound;
    ctl->bound = bound;
    trace_push(a, ctl, 0.0, 0);
    ctl->iter += 1;
    if (ctl->iter > 1 && fabs(bound - ctl->oldbound) < ctl->tol) ctl->converged = 1;
    if (ctl->converged || ctl->iter >= ctl->max_iters || ctl->fault) ctl->done = 1;
    ctl->dg_max = mx;
  }
}

// RCG, after sweep A: Fletcher-Reeves coefficient.  partials (world == 1) or red[K+1] (after all-reduce).
__global__ void rcg_ctl_a_kernel(ViArrays a, ViCtl *ctl, int K, const double *partials, int pstride, int n_ctas) {
  if (ctl->done) return;
  __shared__ double scratch[32];
  double nn;
  if (partials) {
    double acc = 0.0;
    for (int c = threadIdx.x; c < n_ctas; c += CTL_NT) acc += partials[(size_t)c * pstride + K + 1];
    nn = block_sum<CTL_NT>(acc, scratch);
  } else {
    nn = a.red[K + 1];
  }
  if (threadIdx.x == 0) {
    const double beta = nn / ctl->oldnorm;
    ctl->newnorm = nn;
    ctl->oldnorm = nn;
    ctl->beta = beta;
    // the direction memory is empty before the first accepted step (the reference starts it at zero)
    ctl->use_old = (!ctl->didreset && beta > 0.0 && ctl->iter > 0) ? 1 : 0;
    ctl->didreset = 0;
  }
}
// World > 1: bring sweep A's per-CTA partial norms to red[K+1] for the all-reduce.
__global__ void rcg_norm_partial_kernel(ViArrays a, const ViCtl *ctl, int K, const double *partials, int pstride, int n_ctas) {
  if (ctl->done) return;
  __shared__ double scratch[32];
  double acc = 0.0;
  for (int c = threadIdx.x; c < n_ctas; c += CTL_NT) acc += partials[(size_t)c * pstride + K + 1];
  acc = block_sum<CTL_NT>(acc, scratch);
  if (threadIdx.x == 0) a.red[K + 1] = acc;
}

// RCG, after sweep B (stage 0) or after the restart sweep (stage 1).
// red[k] = sum_j c_j q(j,k), red[K] = sum_jk c_j q (logl - gamma).
__global__ void rcg_ctl_b_kernel(ViArrays a, ViCtl *ctl, int K, int stage) {
  if (ctl->done) return;
  if (stage == 1 && !ctl->didreset) return;
  __shared__ double scratch[32];
  __shared__ int s_accept;
  double lg = 0.0;
  for (int k = threadIdx.x; k < K; k += CTL_NT) lg += lgamma(a.alpha0[k] + a.red[k]);
  lg = block_sum<CTL_NT>(lg, scratch);
  const double cand = a.red[K] + lg + ctl->bound_const;
  if (threadIdx.x == 0) {
    if (stage == 0 && cand < ctl->bound) {
      // the conjugate direction lost ground: drop it, redo the step from the same N_k (restart sweep)
      ctl->didreset = 1;
      ctl->resets += 1;
      s_accept = 0;
    } else {
      s_accept = 1;
      ctl->oldbound = ctl->bound;
      ctl->bound = cand;
      trace_push(a, ctl, ctl->newnorm, stage);
      ctl->iter += 1;
      if (stage == 0 && cand - ctl->oldbound < ctl->tol) ctl->converged = 1;
      if (ctl->converged || ctl->iter >= ctl->max_iters) ctl->done = 1;
    }
  }
  __syncthreads();
  if (!s_accept) return;
  for (int k = threadIdx.x; k < K; k += CTL_NT) {
    const double nk = a.alpha0[k] + a.red[k];
    a.N_k[k] = nk;
    a.dg[k] = digamma_series(nk) - 1.0;
  }
}

__global__ void fill_kernel(double *p, size_t n, double v) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
// counts = exp(log_counts); per-block partial sums for the total
__global__ void counts_from_log_kernel(const double *lc, double *c, size_t n, double *block_sums) {
  __shared__ double scratch[32];
  double acc = 0.0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double v = exp(lc[i]);   // exp(-inf) = 0: a class that was not resampled
    c[i] = v;
    acc += v;
  }
  acc = block_sum<256>(acc, scratch);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = acc;
}
__global__ void sum_kernel(const double *c, size_t n, double *block_sums) {
  __shared__ double scratch[32];
  double acc = 0.0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) acc += c[i];
  acc = block_sum<256>(acc, scratch);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = acc;
}
__global__ void sum_blocks_kernel(const double *block_sums, int n, double *out) {
  __shared__ double scratch[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) acc += block_sums[i];
  acc = block_sum<256>(acc, scratch);
  if (threadIdx.x == 0) out[0] = acc;
}

// log-posterior tile in the reference's orientation: out[k * n_rows + j] for rows [row0, row0 + n_rows).
// source 0: stored gamma (RCG).  source 1: recomputed from logl and the last digamma vector (EM).
// source 2: recomputed from P (any storage), rowmax and the last digamma vector.
template <typename ST>
__global__ void posterior_tile_kernel(int source, const double *__restrict__ gamma, const double *__restrict__ logl, int ld,
                                      const ST *__restrict__ P, int ldp, const double *__restrict__ dg,
                                      unsigned long long row0, unsigned long long n_rows, int K, double *__restrict__ out) {
  const int lane = threadIdx.x & 31;
  const unsigned long long warp = (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) >> 5;
  const unsigned long long n_warps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
  for (unsigned long long j = warp; j < n_rows; j += n_warps) {
    const unsigned long long row = row0 + j;
    if (source == 0) {
      for (int k = lane; k < K; k += 32) out[(size_t)k * n_rows + j] = gamma[row * (unsigned long long)ld + k];
      continue;
    }
    double m = -INFINITY;
    for (int k = lane; k < K; k += 32) {
      const double v = source == 1 ? logl[row * (unsigned long long)ld + k] + dg[k]
                                   : log((double)P[row * (unsigned long long)ldp + k]) + dg[k];
      m = fmax(m, v);
    }
    m = warp_max(m);
    double s = 0.0;
    for (int k = lane; k < K; k += 32) {
      const double v = source == 1 ? logl[row * (unsigned long long)ld + k] + dg[k]
                                   : log((double)P[row * (unsigned long long)ldp + k]) + dg[k];
      s += exp(v - m);
    }
    s = warp_sum(s);
    const double lse = m + log(s);
    for (int k = lane; k < K; k += 32) {
      const double v = source == 1 ? logl[row * (unsigned long long)ld + k] + dg[k]
                                   : log((double)P[row * (unsigned long long)ldp + k]) + dg[k];
      out[(size_t)k * n_rows + j] = v - lse;
    }
  }
}

} // namespace mswb

namespace mswb {
// K x n tile of log-posteriors for classes [ec_begin, ec_begin + n) of the shard, on the device (group-major).
void posterior_tile_dev(mswb_ctx *ctx, mswb_lik *lik, uint64_t ec_begin, uint64_t n, double *tile) {
  MSWB_REQUIRE(lik->last_algo >= 0, "no optimisation has been run on this likelihood");
  MSWB_REQUIRE(lik->storage != MSWB_STORE_SPARSE, "posterior export is not available in sparse storage (build the likelihood dense)");
  const int K = (int)lik->K;
  const int blocks = (int)std::min<uint64_t>(ceil_div(n, 8), (uint64_t)ctx->n_sms * 8);
  if (lik->last_algo == MSWB_ALGO_RCG) {
    posterior_tile_kernel<double><<<blocks, 256, 0, ctx->stream>>>(0, lik->gamma.p, nullptr, (int)lik->Kp, nullptr, 0, nullptr, ec_begin, n, K, tile);
  } else if (lik->logl.p) {
    posterior_tile_kernel<double><<<blocks, 256, 0, ctx->stream>>>(1, nullptr, lik->logl.p, (int)lik->Kp, nullptr, 0, lik->last_dg.p, ec_begin, n, K, tile);
  } else if (lik->storage == MSWB_STORE_F32) {
    posterior_tile_kernel<float><<<blocks, 256, 0, ctx->stream>>>(2, nullptr, nullptr, 0, lik->P32.p, (int)lik->Kp32, lik->last_dg.p, ec_begin, n, K, tile);
  } else {
    posterior_tile_kernel<double><<<blocks, 256, 0, ctx->stream>>>(2, nullptr, nullptr, 0, lik->P64.p, (int)lik->Kp, lik->last_dg.p, ec_begin, n, K, tile);
  }
  MSWB_LAUNCHED();
}
} // namespace mswb

// =====================================================================================================
// session
// =====================================================================================================
struct mswb_vi {
  mswb_ctx *ctx = nullptr;
  mswb_lik *lik = nullptr;
  mswb_vi_opts opts{};
  int K = 0;
  bool linear = true;            // EM in the linear domain (P) vs log domain (logl)
  DevBuf<ViCtl> ctl;
  DevBuf<double> alpha0, N_k, dg, dg_prev, w, red, partials, trace_bound, trace_gnorm, own_counts, block_sums;
  DevBuf<unsigned char> trace_reset;
  const double *counts = nullptr;   // device, [N]
  double sum_counts = 0.0;
  std::vector<double> alpha0_host;
  ViArrays arrays{};
  int pstride = 0, grid = 0, max_grid = 0;
  uint64_t enqueued = 0;
  uint64_t pass_bytes = 0;
  // optional kernel timing
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> events;
  size_t events_used = 0;
  double pass_ms_sum = 0.0;
  uint64_t pass_launches = 0;
  ~mswb_vi() { for (auto &e : events) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); } }
};

namespace {

// ---- tile dispatch ---------------------------------------------------------------------------------
// slots = 16-byte pieces per row.  A row is covered by TPR threads x KITER pieces with TPR a multiple of 32 —
// the smallest TPR for the KITER class, so at most one warp's worth of lanes idles whatever K is.  KITER grows with
// the row (1, 2 up to 512 pieces, 4 up to 1024, 8 beyond) and the rows per batch shrink with it:
// R = RMAX / KITER (at most 4), RMAX = 8 for the sweeps that keep two arrays live (EM with its prefetch buffer,
// RCG sweep A), 4 for RCG sweep B (four arrays), so that two CTAs fit an SM (<= 128 registers) up to KITER = 4.
template <int RMAX, int KITER> constexpr int rows_for() { return RMAX / KITER >= 4 ? 4 : (RMAX / KITER >= 2 ? 2 : 1); }

#define MSWB_SHAPE_CASE(TPRV, KITERV, RMAX, ...) { using TL = Tile<TPRV, KITERV, rows_for<RMAX, KITERV>()>; __VA_ARGS__; }
#define MSWB_SHAPE_BY_TPR(tpr, KITERV, RMAX, ...)                                            \
  switch (tpr) {                                                                             \
    case 32: MSWB_SHAPE_CASE(32, KITERV, RMAX, __VA_ARGS__) break;                           \
    case 64: MSWB_SHAPE_CASE(64, KITERV, RMAX, __VA_ARGS__) break;                           \
    case 96: MSWB_SHAPE_CASE(96, KITERV, RMAX, __VA_ARGS__) break;                           \
    case 128: MSWB_SHAPE_CASE(128, KITERV, RMAX, __VA_ARGS__) break;                         \
    case 160: MSWB_SHAPE_CASE(160, KITERV, RMAX, __VA_ARGS__) break;                         \
    case 192: MSWB_SHAPE_CASE(192, KITERV, RMAX, __VA_ARGS__) break;                         \
    case 224: MSWB_SHAPE_CASE(224, KITERV, RMAX, __VA_ARGS__) break;                         \
    default: MSWB_SHAPE_CASE(256, KITERV, RMAX, __VA_ARGS__) break;                          \
  }
// rows of more than 512 pieces need at least 160 threads per row in the KITER = 4 / 8 classes
#define MSWB_SHAPE_BY_TPR_WIDE(tpr, KITERV, RMAX, ...)                                       \
  switch (tpr) {                                                                             \
    case 160: MSWB_SHAPE_CASE(160, KITERV, RMAX, __VA_ARGS__) break;                         \
    case 192: MSWB_SHAPE_CASE(192, KITERV, RMAX, __VA_ARGS__) break;                         \
    case 224: MSWB_SHAPE_CASE(224, KITERV, RMAX, __VA_ARGS__) break;                         \
    default: MSWB_SHAPE_CASE(256, KITERV, RMAX, __VA_ARGS__) break;                          \
  }
// fp32 rows that would need 224 / 256 threads with more registers than two CTAs per SM allow (7-8 warps per SM: 0.75 of
// the peak at K = 4000) take a 512-thread CTA with half the pieces instead (1.00).  fp64 keeps 256 x 8 (0.83 vs 0.74).
#define MSWB_TILE_DISPATCH(slots, RMAX, IS_F32, ...)                                         \
  do {                                                                                       \
    const int _s = (int)(slots);                                                             \
    if (_s <= 32) MSWB_SHAPE_CASE(32, 1, RMAX, __VA_ARGS__)                                  \
    else if (_s <= 512) { const int _t = (int)round_up(ceil_div(_s, 2), 32); MSWB_SHAPE_BY_TPR(_t, 2, RMAX, __VA_ARGS__) }   \
    else if (_s <= 1024) { const int _t = (int)round_up(ceil_div(_s, 4), 32);                \
      if ((IS_F32) && _t >= 224) MSWB_SHAPE_CASE(512, 2, RMAX, __VA_ARGS__)                  \
      else MSWB_SHAPE_BY_TPR_WIDE(_t, 4, RMAX, __VA_ARGS__) }                                \
    else if (_s <= 2048) { const int _t = (int)round_up(ceil_div(_s, 8), 32);                \
      if ((IS_F32) && _t >= 224) MSWB_SHAPE_CASE(512, 4, 4, __VA_ARGS__)                     \
      else MSWB_SHAPE_BY_TPR_WIDE(_t, 8, RMAX, __VA_ARGS__) }                                \
    else if (_s <= 4096) MSWB_SHAPE_CASE(512, 8, RMAX, __VA_ARGS__)                          \
    else if (_s <= 8192) MSWB_SHAPE_CASE(1024, 8, RMAX, __VA_ARGS__)                         \
    else throw Error("too many groups for the compiled tile shapes (max 16384 in fp64)");   \
  } while (0)

// Log-domain (RCG) sweeps.  Rows of 64-256 threads are fed by the two-CTA TMA ring (launch_sweep_a / _b), which made
// lane efficiency count: three pieces per thread and the 96 / 192 / 224-thread rows below each gained 4-20 % over the
// next power-of-two shape once the ring was in (before it, with direct loads, they lost 3-7 %: fewer bytes in flight
// per thread).  160-thread rows lost (K = 900: 0.85 -> 0.66 of the peak: five-warp CTAs, two per SM) and are not used.
// One-warp rows (<= 128 pieces) keep direct loads.  1025-2048 pieces (K = 2050..4096): 512 threads x 4 pieces
// (126 registers, 16 warps per SM); 256 threads x 8 pieces needs ~250 registers in sweep B and ran at 0.40 instead of
// 0.60-0.77 of the peak.
#define MSWB_TILE_DISPATCH_RCG(slots, RMAX, ...)                                             \
  do {                                                                                       \
    const int _s = (int)(slots);                                                             \
    if (_s <= 32) MSWB_SHAPE_CASE(32, 1, RMAX, __VA_ARGS__)                                  \
    else if (_s <= 64) MSWB_SHAPE_CASE(32, 2, RMAX, __VA_ARGS__)                             \
    else if (_s <= 128) MSWB_SHAPE_CASE(32, 4, RMAX, __VA_ARGS__)                            \
    else if (_s <= 192) MSWB_SHAPE_CASE(64, 3, RMAX, __VA_ARGS__)                            \
    else if (_s <= 256) MSWB_SHAPE_CASE(64, 4, RMAX, __VA_ARGS__)                            \
    else if (_s <= 288) MSWB_SHAPE_CASE(96, 3, RMAX, __VA_ARGS__)                            \
    else if (_s <= 384) MSWB_SHAPE_CASE(128, 3, RMAX, __VA_ARGS__)                           \
    else if (_s <= 512) MSWB_SHAPE_CASE(128, 4, RMAX, __VA_ARGS__)                           \
    else if (_s <= 576) MSWB_SHAPE_CASE(192, 3, RMAX, __VA_ARGS__)                           \
    else if (_s > 640 && _s <= 672) MSWB_SHAPE_CASE(224, 3, RMAX, __VA_ARGS__)               \
    else if (_s <= 768) MSWB_SHAPE_CASE(256, 3, RMAX, __VA_ARGS__)                           \
    else if (_s <= 896) MSWB_SHAPE_CASE(224, 4, RMAX, __VA_ARGS__)                           \
    else if (_s <= 1024) MSWB_SHAPE_CASE(256, 4, RMAX, __VA_ARGS__)                          \
    else if (_s <= 2048) MSWB_SHAPE_CASE(512, 4, RMAX, __VA_ARGS__)                          \
    else if (_s <= 4096) MSWB_SHAPE_CASE(512, 8, RMAX, __VA_ARGS__)                          \
    else if (_s <= 8192) MSWB_SHAPE_CASE(1024, 8, RMAX, __VA_ARGS__)                         \
    else throw Error("too many groups for the compiled tile shapes (max 16384 in fp64)");   \
  } while (0)

constexpr size_t SMEM_BUDGET = 200 * 1024;   // dynamic shared memory for the stage ring (of 227 KB per CTA)

// Stage ring geometry for a sweep that streams `nsrc` arrays with rows of `row_bytes`, consumed in
// units of `unit_rows` (= G*R).  stages == 0: the rows are too long to stage, use the direct kernel.
PipeGeom pipe_geometry(size_t row_bytes, int unit_rows, int nsrc, int tpr, size_t budget = SMEM_BUDGET, size_t target = 32 * 1024) {
  PipeGeom g{0, 0, 0};
  if (tpr > 512) return g;
  const size_t unit_bytes = (size_t)unit_rows * row_bytes;
  if (const char *e = getenv("MSWB_STAGE_KB")) target = (size_t)atoi(e) * 1024;
  const int units = (int)std::max<size_t>(1, target / unit_bytes);
  g.stage_rows = unit_rows * units;
  g.stage_pitch = (unsigned)round_up((size_t)g.stage_rows * row_bytes, 128);
  const size_t per_stage = (size_t)nsrc * g.stage_pitch + 8;
  if (const char *e = getenv("MSWB_SMEM_KB")) budget = (size_t)atoi(e) * 1024;
  int stages = (int)std::min<size_t>(8, budget / per_stage);
  if (const char *e = getenv("MSWB_STAGES")) stages = std::min(stages, atoi(e));
  g.stages = stages >= 2 ? stages : 0;
  return g;
}
size_t pipe_smem_bytes(const PipeGeom &g, int nsrc) { return (size_t)g.stages * nsrc * g.stage_pitch + (size_t)g.stages * 8; }

// Persistent grid: as many CTAs as are resident at once (one per SM for the staged kernels).  The occupancy query
// is cached per (kernel, block size, shared memory): small problems run thousands of launches per second.
template <class Kern> int persistent_grid(mswb_ctx *ctx, Kern kern, int nt, size_t smem, uint64_t n_batches, int max_grid) {
  struct Key { const void *k; int nt; size_t smem; bool operator<(const Key &o) const { return std::tie(k, nt, smem) < std::tie(o.k, o.nt, o.smem); } };
  static std::map<Key, int> cache;
  static std::mutex mu;
  int per_sm;
  {
    std::lock_guard<std::mutex> lock(mu);
    const Key key{(const void *)kern, nt, smem};
    auto it = cache.find(key);
    if (it == cache.end()) {
      if (smem > 48 * 1024) MSWB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int v = 1;
      MSWB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kern, nt, smem));
      it = cache.emplace(key, v < 1 ? 1 : v).first;
    }
    per_sm = it->second;
  }
  uint64_t g = (uint64_t)ctx->n_sms * per_sm;
  if (g > n_batches) g = n_batches;
  if (g > (uint64_t)max_grid) g = max_grid;
  if (g < 1) g = 1;
  return (int)g;
}

struct PassTimer {
  mswb_vi *vi; bool on; size_t slot = 0;
  explicit PassTimer(mswb_vi *v) : vi(v), on(v->opts.time_kernels != 0) {
    if (!on) return;
    if (vi->events_used == vi->events.size()) {
      cudaEvent_t a, b;
      MSWB_CUDA(cudaEventCreate(&a)); MSWB_CUDA(cudaEventCreate(&b));
      vi->events.emplace_back(a, b);
    }
    slot = vi->events_used++;
    MSWB_CUDA(cudaEventRecord(vi->events[slot].first, vi->ctx->stream));
  }
  void stop() { if (on) MSWB_CUDA(cudaEventRecord(vi->events[slot].second, vi->ctx->stream)); }
};

void launch_finalize(mswb_vi *vi, int nvals, int only_if_reset) {
  const int nt = 128;
  finalize_partials_kernel<<<(nvals + nt - 1) / nt, nt, 0, vi->ctx->stream>>>(
      vi->partials.p, vi->pstride, vi->grid, nvals, vi->red.p, vi->ctl.p, only_if_reset);
  MSWB_LAUNCHED();
}

// Measured on B200 (1e6 x 2000 fp64).  A ring that takes the whole SM (200 KB, one CTA) loses to direct streaming
// loads with two CTAs per SM: log-domain sweeps 4.0-4.3 vs 5.5 TB/s, EM sweep 5.2 vs 6.5-6.7 TB/s — one CTA marches in
// lockstep through load / exp / reduce phases.  A ring sized for TWO resident CTAs (100 KB, 16 KB stages: double
// buffering for sweep B, three stages for sweep A) beats both for the exp-heavy sweeps — bytes in flight no longer
// depend on registers: K = 2000 5.68 vs 5.46 TB/s, K = 1500 5.4 vs 4.6, K = 1000 5.75 vs 5.44, K = 700 4.68 vs 4.39,
// K = 420 5.16 vs 4.93.  The 512-thread rows (K = 2050..4096) hold one CTA per SM either way and take the ring with the
// whole 200 KB (K = 3000: 4.35 vs 3.84 TB/s).  Of the one-warp rows (TPR = 32, K <= 256) only the two-piece shape
// (33-64 pieces, K = 66..128) takes the ring: K = 70 +8 %, K = 100 +6 %, K = 128 -1 %; with one or four pieces per
// thread it gained nothing consistent (K = 50 +1 %, K = 150 -1 %, K = 256 -4 %) and those keep direct loads.  So: RCG sweeps with rows of 64-256 threads use the two-CTA ring by default (MSWB_RCG_TMA=0 turns
// it off); the EM sweep, already at the copy peak with direct loads, keeps them (MSWB_EM_TMA=1 selects the one-CTA ring,
// compiled for TPR = 256 only).
constexpr size_t RCG_RING_BYTES = 100 * 1024, RCG_STAGE_BYTES = 16 * 1024;
bool want_rcg_pipe() { const char *e = getenv("MSWB_RCG_TMA"); return !(e && e[0] == '0'); }
bool want_em_pipe() { const char *e = getenv("MSWB_EM_TMA"); return e && e[0] == '1'; }

// Batch -> CTA mapping of the direct EM sweep (see the kernel): chunked once the matrix is large.
static bool em_chunked(const mswb_lik *L) {
  if (const char *e = getenv("MSWB_EM_CHUNKED")) return e[0] == '1';
  const size_t el = L->storage == MSWB_STORE_F32 ? 4 : 8;
  return (size_t)L->N_pad * L->K * el > ((size_t)16 << 30);   // measured: +1.2 % at 100 GB, neutral at 24 GB
}

template <typename ST, class TL> void launch_em(mswb_vi *vi, const ST *P, int ld) {
  mswb_lik *L = vi->lik;
  cudaStream_t s = vi->ctx->stream;
  PipeGeom geom{0, 0, 0};
  if constexpr (TL::TPR == 256) {
    if (want_em_pipe()) geom = pipe_geometry((size_t)ld * sizeof(ST), TL::G * TL::R, 1, TL::TPR);
    if (geom.stages) {
      auto kern = em_lin_pass_kernel<ST, TL, true>;
      const size_t smem = pipe_smem_bytes(geom, 1);
      vi->grid = persistent_grid(vi->ctx, kern, TL::NT, smem, ceil_div(L->N_pad, (uint64_t)geom.stage_rows), vi->max_grid);
      kern<<<vi->grid, TL::NT, smem, s>>>(P, ld, L->rowmax.p, vi->counts, vi->w.p, vi->ctl.p, vi->partials.p, vi->pstride, L->N_pad,
                                          vi->K, geom);
      MSWB_LAUNCHED();
      return;
    }
  }
  if (em_chunked(L)) geom.stage_rows = 1;
  auto kern = em_lin_pass_kernel<ST, TL, false>;
  vi->grid = persistent_grid(vi->ctx, kern, TL::NT, 0, L->N_pad / (TL::G * TL::R), vi->max_grid);
  kern<<<vi->grid, TL::NT, 0, s>>>(P, ld, L->rowmax.p, vi->counts, vi->w.p, vi->ctl.p, vi->partials.p, vi->pstride, L->N_pad, vi->K, geom);
  MSWB_LAUNCHED();
}

template <class TL> void launch_sweep_a(mswb_vi *vi) {
  mswb_lik *L = vi->lik;
  cudaStream_t s = vi->ctx->stream;
  const int ld = (int)L->Kp;
  PipeGeom geom{0, 0, 0};
  if constexpr ((TL::TPR >= 64 || (TL::TPR == 32 && TL::KITER == 2)) && TL::NT <= 512) {
    if (want_rcg_pipe()) geom = pipe_geometry((size_t)ld * 8, TL::G * TL::R, 2, TL::TPR, TL::NT <= 256 ? RCG_RING_BYTES : SMEM_BUDGET, TL::NT <= 256 ? RCG_STAGE_BYTES : 2 * RCG_STAGE_BYTES);
    if (geom.stages) {
      auto kern = rcg_sweep_a_kernel<TL, true>;
      const size_t smem = pipe_smem_bytes(geom, 2);
      vi->grid = persistent_grid(vi->ctx, kern, TL::NT, smem, ceil_div(L->N, (uint64_t)geom.stage_rows), vi->max_grid);
      kern<<<vi->grid, TL::NT, smem, s>>>(L->logl.p, L->gamma.p, ld, vi->dg.p, vi->ctl.p, vi->partials.p, vi->pstride, L->N, vi->K, geom);
      MSWB_LAUNCHED();
      return;
    }
  }
  auto kern = rcg_sweep_a_kernel<TL, false>;
  vi->grid = persistent_grid(vi->ctx, kern, TL::NT, 0, ceil_div(L->N, (uint64_t)TL::G * TL::R), vi->max_grid);
  kern<<<vi->grid, TL::NT, 0, s>>>(L->logl.p, L->gamma.p, ld, vi->dg.p, vi->ctl.p, vi->partials.p, vi->pstride, L->N, vi->K, geom);
  MSWB_LAUNCHED();
}

template <class TL, int MODE, bool WRITE> void launch_sweep_b(mswb_vi *vi, int only_if_reset) {
  mswb_lik *L = vi->lik;
  cudaStream_t s = vi->ctx->stream;
  const int ld = (int)L->Kp;
  double *gam = WRITE || MODE == 0 ? L->gamma.p : nullptr, *stp = MODE == 0 ? L->step.p : nullptr;
  PipeGeom geom{0, 0, 0};
  if constexpr ((TL::TPR >= 64 || (TL::TPR == 32 && TL::KITER == 2)) && TL::NT <= 512) {
    if (want_rcg_pipe()) geom = pipe_geometry((size_t)ld * 8, TL::G * TL::R, 3, TL::TPR, TL::NT <= 256 ? RCG_RING_BYTES : SMEM_BUDGET, TL::NT <= 256 ? RCG_STAGE_BYTES : 2 * RCG_STAGE_BYTES);
    if (geom.stages) {
      auto kern = rcg_sweep_b_kernel<TL, MODE, WRITE, true>;
      const size_t smem = pipe_smem_bytes(geom, 3);
      vi->grid = persistent_grid(vi->ctx, kern, TL::NT, smem, ceil_div(L->N, (uint64_t)geom.stage_rows), vi->max_grid);
      kern<<<vi->grid, TL::NT, smem, s>>>(L->logl.p, gam, stp, ld, vi->dg.p, vi->counts, vi->ctl.p, vi->partials.p, vi->pstride, L->N,
                                          vi->K, only_if_reset, geom);
      MSWB_LAUNCHED();
      return;
    }
  }
  auto kern = rcg_sweep_b_kernel<TL, MODE, WRITE, false>;
  vi->grid = persistent_grid(vi->ctx, kern, TL::NT, 0, ceil_div(L->N, (uint64_t)TL::G * TL::R), vi->max_grid);
  kern<<<vi->grid, TL::NT, 0, s>>>(L->logl.p, gam, stp, ld, vi->dg.p, vi->counts, vi->ctl.p, vi->partials.p, vi->pstride, L->N, vi->K,
                                   only_if_reset, geom);
  MSWB_LAUNCHED();
}

void em_iteration(mswb_vi *vi) {
  mswb_lik *L = vi->lik;
  cudaStream_t s = vi->ctx->stream;
  const int K = vi->K;
  PassTimer timer(vi);
  if (vi->linear && L->storage == MSWB_STORE_SPARSE) {
    const size_t smem = (size_t)2 * K * sizeof(double);
    MSWB_REQUIRE(smem <= 200 * 1024, "too many groups for the sparse EM pass (weights and accumulators live in shared memory)");
    if (smem > 48 * 1024) MSWB_CUDA(cudaFuncSetAttribute(em_sparse_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    MSWB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, em_sparse_pass_kernel, 256, smem));
    vi->grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)vi->ctx->n_sms * std::max(1, per_sm), std::min<uint64_t>(ceil_div(L->N, 256), (uint64_t)vi->max_grid)));
    em_sparse_pass_kernel<<<vi->grid, 256, smem, s>>>(L->nz_ptr.p, L->nz_grp.p, L->nz_dP.p, L->P0.p, L->rowmax.p, vi->counts, vi->w.p,
                                                     vi->ctl.p, vi->partials.p, vi->pstride, L->N, K);
    MSWB_LAUNCHED();
  } else if (vi->linear && L->storage == MSWB_STORE_F32) {
    MSWB_TILE_DISPATCH(L->Kp32 / 4, 8, true, launch_em<float, TL>(vi, L->P32.p, (int)L->Kp32));
  } else if (vi->linear) {
    MSWB_TILE_DISPATCH(L->Kp / 2, 8, false, launch_em<double, TL>(vi, L->P64.p, (int)L->Kp));
  } else {
    MSWB_TILE_DISPATCH_RCG(L->Kp / 2, 4, launch_sweep_b<TL, 1, false>(vi, 0));
  }
  timer.stop();
  launch_finalize(vi, vi->linear ? K + 2 : K + 1, 0);
  vi->ctx->allreduce_sum(vi->red.p, vi->linear ? K + 2 : K + 1);
  em_ctl_kernel<<<1, CTL_NT, 0, s>>>(vi->arrays, vi->ctl.p, K, vi->linear ? 1 : 0);
  MSWB_LAUNCHED();
}

void rcg_iteration(mswb_vi *vi) {
  mswb_lik *L = vi->lik;
  mswb_ctx *ctx = vi->ctx;
  cudaStream_t s = ctx->stream;
  const int K = vi->K;
  const int slots = L->Kp / 2;
  // sweep A: gradient norm
  {
    PassTimer timer(vi);
    MSWB_TILE_DISPATCH_RCG(slots, 8, launch_sweep_a<TL>(vi));   // (one row per batch / three CTAs per SM measured the same: 5456 vs 5445 GB/s)
    timer.stop();
  }
  if (ctx->world > 1) {
    rcg_norm_partial_kernel<<<1, CTL_NT, 0, s>>>(vi->arrays, vi->ctl.p, K, vi->partials.p, vi->pstride, vi->grid);
    MSWB_LAUNCHED();
    ctx->allreduce_sum(vi->red.p + K + 1, 1);
    rcg_ctl_a_kernel<<<1, CTL_NT, 0, s>>>(vi->arrays, vi->ctl.p, K, nullptr, 0, 0);
  } else {
    rcg_ctl_a_kernel<<<1, CTL_NT, 0, s>>>(vi->arrays, vi->ctl.p, K, vi->partials.p, vi->pstride, vi->grid);
  }
  MSWB_LAUNCHED();
  // sweep B: step, renormalise, N_k, bound
  {
    PassTimer timer(vi);
    MSWB_TILE_DISPATCH_RCG(slots, 4, launch_sweep_b<TL, 0, true>(vi, 0));
    timer.stop();
  }
  launch_finalize(vi, K + 1, 0);
  ctx->allreduce_sum(vi->red.p, K + 1);
  rcg_ctl_b_kernel<<<1, CTL_NT, 0, s>>>(vi->arrays, vi->ctl.p, K, 0);
  MSWB_LAUNCHED();
  // restart sweep: runs only when the control block says so (device-side decision, no host round trip)
  MSWB_TILE_DISPATCH_RCG(slots, 4, launch_sweep_b<TL, 1, true>(vi, 1));
  launch_finalize(vi, K + 1, 1);
  ctx->allreduce_sum(vi->red.p, K + 1);
  rcg_ctl_b_kernel<<<1, CTL_NT, 0, s>>>(vi->arrays, vi->ctl.p, K, 1);
  MSWB_LAUNCHED();
}

void fill_stat(mswb_vi *vi, const ViCtl &c, mswb_vi_stat *stat) {
  if (!stat) return;
  stat->bound = c.bound;
  stat->gnorm = c.newnorm;
  stat->iters = c.iter;
  stat->converged = c.converged;
  stat->resets = c.resets;
  stat->pass_ms_sum = vi->pass_ms_sum;
  stat->pass_launches = vi->pass_launches;
  stat->pass_bytes = vi->pass_bytes;
}

ViCtl poll_ctl(mswb_vi *vi) {
  ViCtl c;
  d2h(&c, vi->ctl.p, 1, vi->ctx->stream);
  MSWB_CUDA(cudaStreamSynchronize(vi->ctx->stream));
  if (vi->opts.time_kernels) {
    for (size_t i = 0; i < vi->events_used; ++i) {
      float ms = 0.f;
      MSWB_CUDA(cudaEventElapsedTime(&ms, vi->events[i].first, vi->events[i].second));
      vi->pass_ms_sum += ms;
      vi->pass_launches += 1;
    }
    vi->events_used = 0;
  }
  MSWB_REQUIRE(!c.fault, "EM pass: a class normaliser under/overflowed in the linear domain (extreme prior counts)");
  return c;
}

} // namespace

// counts_dev != NULL: class counts already on the device with their (global) total `counts_dev_sum`.
static int vi_begin_impl(mswb_ctx *ctx, mswb_lik *lik, const double *alpha0, const double *log_counts,
                         const double *counts_dev, double counts_dev_sum, const mswb_vi_opts *opts, mswb_vi **out) {
  return guarded([&] {
    MSWB_REQUIRE(ctx && lik && alpha0 && opts && out, "NULL argument");
    MSWB_REQUIRE(lik->ctx == ctx, "likelihood belongs to another context");
    MSWB_REQUIRE(opts->algo == MSWB_ALGO_RCG || opts->algo == MSWB_ALGO_EM, "unknown algorithm");
    MSWB_REQUIRE(lik->K >= 1, "likelihood has no groups");
    MSWB_CUDA(cudaSetDevice(ctx->device));
    std::unique_ptr<mswb_vi> vi(new mswb_vi);
    vi->ctx = ctx; vi->lik = lik; vi->opts = *opts;
    const int K = vi->K = (int)lik->K;
    cudaStream_t s = ctx->stream;

    if (opts->algo == MSWB_ALGO_RCG) {
      MSWB_REQUIRE(lik->storage == MSWB_STORE_F64, "RCG needs the fp64 log-likelihood (build the likelihood with MSWB_STORE_F64)");
      lik_ensure_logl(lik);
      const size_t n = (size_t)lik->N * lik->Kp;
      lik->gamma.ensure(n);
      lik->step.ensure(n);
      fill_kernel<<<ctx->n_sms * 4, 256, 0, s>>>(lik->gamma.p, n, std::log(1.0 / (double)K));
      MSWB_LAUNCHED();
      vi->pass_bytes = (uint64_t)lik->N * K * 56 + (uint64_t)lik->N * 8;   // sweep A 16 B + sweep B 40 B per element
    } else {
      vi->linear = true;
      if (lik->storage == MSWB_STORE_SPARSE) {
        lik_ensure_sparse(lik);
        vi->pass_bytes = lik->nnz * 12 + (uint64_t)lik->N * 40;              // hits (group + value), ptr, P0, M_j, c_j
      } else {
        lik_ensure_linear(lik);
        const uint64_t bl = lik->storage == MSWB_STORE_F32 ? 4 : 8;
        vi->pass_bytes = (uint64_t)lik->N * K * bl + (uint64_t)lik->N * 16;  // P once, c_j and M_j once
      }
    }

    // per-group vectors
    vi->alpha0.alloc(K); vi->N_k.alloc(K); vi->dg.alloc(K); vi->dg_prev.alloc(K); vi->w.alloc(K); vi->red.alloc(K + 2);
    vi->alpha0_host.assign(alpha0, alpha0 + K);
    for (int k = 0; k < K; ++k) MSWB_REQUIRE(alpha0[k] > 0.0 && std::isfinite(alpha0[k]), "prior counts must be positive");
    h2d(vi->alpha0.p, alpha0, K, s);
    vi->max_grid = ctx->n_sms * 8;
    vi->pstride = (int)round_up(K + 2, 2);
    vi->partials.alloc((size_t)vi->max_grid * vi->pstride);
    const uint64_t cap = std::min<uint64_t>(opts->max_iters, 1u << 20);
    vi->trace_bound.alloc(cap); vi->trace_gnorm.alloc(cap); vi->trace_reset.alloc(cap);
    vi->ctl.alloc(1);

    // class counts
    if (counts_dev) {
      vi->counts = counts_dev;
      vi->sum_counts = counts_dev_sum;
    } else if (log_counts) {
      DevBuf<double> lc;
      lc.alloc(lik->N);
      h2d(lc.p, log_counts, lik->N, s);
      vi->own_counts.alloc(lik->N_pad);
      MSWB_CUDA(cudaMemsetAsync(vi->own_counts.p, 0, vi->own_counts.bytes(), s));
      const int nb = 296;
      vi->block_sums.ensure(nb + 1);
      counts_from_log_kernel<<<nb, 256, 0, s>>>(lc.p, vi->own_counts.p, lik->N, vi->block_sums.p);
      MSWB_LAUNCHED();
      sum_blocks_kernel<<<1, 256, 0, s>>>(vi->block_sums.p, nb, vi->block_sums.p + nb);
      MSWB_LAUNCHED();
      ctx->allreduce_sum(vi->block_sums.p + nb, 1);
      d2h(&vi->sum_counts, vi->block_sums.p + nb, 1, s);
      MSWB_CUDA(cudaStreamSynchronize(s));
      vi->counts = vi->own_counts.p;
    } else {
      vi->counts = lik->counts.p;
      vi->sum_counts = lik->sum_counts_total;
    }

    // bound constant: lgamma(sum alpha0) - lgamma(sum alpha0 + sum c) - sum lgamma(alpha0)
    long double a_sum = 0.0L, lg_sum = 0.0L;
    for (int k = 0; k < K; ++k) { a_sum += alpha0[k]; lg_sum += std::lgamma(alpha0[k]); }
    const double bconst = (double)(std::lgamma((double)a_sum) - std::lgamma((double)(a_sum + vi->sum_counts)) - lg_sum);

    vi->arrays = ViArrays{vi->alpha0.p, vi->N_k.p, vi->dg.p, vi->w.p, vi->dg_prev.p, vi->red.p, vi->trace_bound.p,
                          vi->trace_gnorm.p, vi->trace_reset.p, cap};
    vi_init_kernel<<<1, CTL_NT, 0, s>>>(vi->arrays, vi->ctl.p, K, opts->algo, opts->tol, opts->max_iters, bconst, vi->sum_counts);
    MSWB_LAUNCHED();
    lik->last_algo = opts->algo;
    *out = vi.release();
  });
}

extern "C" {

int mswb_vi_begin(mswb_ctx *ctx, mswb_lik *lik, const double *alpha0, const double *log_counts,
                  const mswb_vi_opts *opts, mswb_vi **out) {
  return vi_begin_impl(ctx, lik, alpha0, log_counts, nullptr, 0.0, opts, out);
}

int mswb_vi_step(mswb_vi *vi, uint64_t n_iters) {
  return guarded([&] {
    MSWB_REQUIRE(vi, "vi is NULL");
    MSWB_CUDA(cudaSetDevice(vi->ctx->device));
    for (uint64_t i = 0; i < n_iters; ++i) {
      if (vi->opts.algo == MSWB_ALGO_RCG) rcg_iteration(vi); else em_iteration(vi);
    }
    vi->enqueued += n_iters;
  });
}

int mswb_vi_poll(mswb_vi *vi, mswb_vi_stat *stat) {
  return guarded([&] {
    MSWB_REQUIRE(vi, "vi is NULL");
    MSWB_CUDA(cudaSetDevice(vi->ctx->device));
    const ViCtl c = poll_ctl(vi);
    fill_stat(vi, c, stat);
  });
}

int mswb_vi_trace(mswb_vi *vi, double *bound, double *gnorm, uint8_t *reset, uint64_t capacity) {
  return guarded([&] {
    MSWB_REQUIRE(vi, "vi is NULL");
    MSWB_CUDA(cudaSetDevice(vi->ctx->device));
    const ViCtl c = poll_ctl(vi);
    const uint64_t n = std::min<uint64_t>(std::min<uint64_t>(c.iter, capacity), vi->arrays.trace_cap);
    cudaStream_t s = vi->ctx->stream;
    if (bound) d2h(bound, vi->trace_bound.p, n, s);
    if (gnorm) d2h(gnorm, vi->trace_gnorm.p, n, s);
    if (reset) d2h(reset, vi->trace_reset.p, n, s);
    MSWB_CUDA(cudaStreamSynchronize(s));
  });
}

int mswb_vi_finish(mswb_vi *vi, double *theta, double *N_k, mswb_vi_stat *stat) {
  if (!vi) { set_last_error("vi is NULL"); return 1; }
  std::unique_ptr<mswb_vi> own(vi);
  return guarded([&] {
    MSWB_CUDA(cudaSetDevice(vi->ctx->device));
    const ViCtl c = poll_ctl(vi);
    std::vector<double> nk(vi->K);
    d2h(nk.data(), vi->N_k.p, vi->K, vi->ctx->stream);
    if (vi->opts.algo == MSWB_ALGO_EM) {
      // gamma is not stored for EM: keep the digamma vector the last pass used; it reproduces the
      // responsibilities of that pass on demand (mswb_vi_posteriors).
      vi->lik->last_dg.ensure(vi->K);
      MSWB_CUDA(cudaMemcpyAsync(vi->lik->last_dg.p, c.iter > 0 ? vi->dg_prev.p : vi->dg.p, vi->K * sizeof(double), cudaMemcpyDeviceToDevice, vi->ctx->stream));
    }
    MSWB_CUDA(cudaStreamSynchronize(vi->ctx->stream));
    // rcgpar::mixture_components: theta_k = sum_j c_j q(j,k) / sum_j c_j = (N_k - alpha0_k) / sum_j c_j
    if (theta) for (int k = 0; k < vi->K; ++k) theta[k] = (nk[k] - vi->alpha0_host[k]) / vi->sum_counts;
    if (N_k) std::memcpy(N_k, nk.data(), vi->K * sizeof(double));
    fill_stat(vi, c, stat);
  });
}

static int vi_run_impl(mswb_ctx *ctx, mswb_lik *lik, const double *alpha0, const double *log_counts,
                       const double *counts_dev, double counts_dev_sum, const mswb_vi_opts *opts, double *theta,
                       mswb_vi_stat *stat, mswb_iter_cb on_iter, void *user) {
  mswb_vi *vi = nullptr;
  if (vi_begin_impl(ctx, lik, alpha0, log_counts, counts_dev, counts_dev_sum, opts, &vi)) return 1;
  int rc = guarded([&] {
    const uint64_t every = opts->poll_every ? opts->poll_every : 8;
    uint64_t reported = 0;
    std::vector<double> tb, tg;
    for (;;) {
      if (mswb_vi_step(vi, every)) throw Error(mswb_last_error());
      const ViCtl c = poll_ctl(vi);
      if (on_iter && c.iter > reported) {
        const uint64_t n = std::min<uint64_t>(c.iter, vi->arrays.trace_cap);
        tb.resize(n); tg.resize(n);
        d2h(tb.data() + reported, vi->trace_bound.p + reported, n - reported, ctx->stream);
        d2h(tg.data() + reported, vi->trace_gnorm.p + reported, n - reported, ctx->stream);
        MSWB_CUDA(cudaStreamSynchronize(ctx->stream));
        for (uint64_t i = reported; i < n; ++i) on_iter(user, i, tb[i], tg[i]);
        reported = n;
      }
      if (c.done) break;
    }
  });
  if (rc) { std::string keep = mswb_last_error(); mswb_vi_finish(vi, nullptr, nullptr, nullptr); set_last_error(keep); return 1; }
  return mswb_vi_finish(vi, theta, nullptr, stat);
}

int mswb_vi_run(mswb_ctx *ctx, mswb_lik *lik, const double *alpha0, const double *log_counts,
                const mswb_vi_opts *opts, double *theta, mswb_vi_stat *stat, mswb_iter_cb on_iter, void *user) {
  return vi_run_impl(ctx, lik, alpha0, log_counts, nullptr, 0.0, opts, theta, stat, on_iter, user);
}

int mswb_vi_posteriors(mswb_ctx *ctx, mswb_lik *lik, uint64_t ec_begin, uint64_t ec_end, double *gamma) {
  return guarded([&] {
    MSWB_REQUIRE(ctx && lik && gamma, "NULL argument");
    MSWB_REQUIRE(ec_begin <= ec_end && ec_end <= lik->N, "class range out of bounds");
    MSWB_CUDA(cudaSetDevice(ctx->device));
    const uint64_t n = ec_end - ec_begin;
    if (n == 0) return;
    DevBuf<double> tile;
    tile.alloc((size_t)n * lik->K);
    mswb::posterior_tile_dev(ctx, lik, ec_begin, n, tile.p);
    d2h(gamma, tile.p, (size_t)n * lik->K, ctx->stream);
    MSWB_CUDA(cudaStreamSynchronize(ctx->stream));
  });
}

} // extern "C"

// internal (bootstrap.cu): class counts already resident on the device
int mswb_vi_run_dev_counts(mswb_ctx *ctx, mswb_lik *lik, const double *alpha0, const double *counts_dev,
                           double sum_counts, const mswb_vi_opts *opts, double *theta, mswb_vi_stat *stat) {
  return vi_run_impl(ctx, lik, alpha0, nullptr, counts_dev, sum_counts, opts, theta, stat, nullptr, nullptr);
}
