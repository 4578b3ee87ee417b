// vi.cu — variational optimiser: control kernels, tile dispatch and the host-side session.
// Replaces rcgpar::rcg_optl_omp / rcg_optl_torch / em_torch / mixture_components as called from the
// reference at src/mSWEEP.cpp:192-203, 419-423, 507-516.
#include "handles.cuh"
#include "vi_kernels.cuh"
#include "vi_sparse_rcg.cuh"

#include <cmath>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <mutex>
#include <tuple>

using namespace mswb;

// =====================================================================================================
// control kernels (one CTA; K-sized work).  The control STEPS live in vi_kernels.cuh: on one GPU they run in the
// tail of the sweeps (or of finalize_ctl_kernel); these wrappers serve several GPUs, where the all-reduce sits between.
// =====================================================================================================
namespace mswb {

constexpr int CTL_NT = 256;

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
    ctl->stall = 0; ctl->ticket = 0u; ctl->epoch = 0u;
    ctl->done = max_iters == 0 ? 1 : 0;
  }
}

__global__ void __launch_bounds__(CTL_NT) em_ctl_kernel(ViArrays a, ViCtl *ctl, int K, int sparse) {
  if (ctl->done) return;
  __shared__ double scratch[32];
  em_ctl_step<CTL_NT>(a, ctl, K, sparse, scratch);
}

// stage 0: after sweep B.  stage 1: after the restart sweep the host enqueued for a stalled optimisation.
__global__ void __launch_bounds__(CTL_NT) rcg_ctl_b_kernel(ViArrays a, ViCtl *ctl, int K, int stage, int stall_on_reject) {
  if (ctl->done) return;
  if (stage == 0 ? ctl->stall != 0 : !ctl->didreset) return;
  __shared__ double scratch[32];
  rcg_ctl_b_step<CTL_NT>(a, ctl, K, stage, stall_on_reject, scratch);
}

// Several GPUs with peer memory (peer.cuh): the exchange of the reduced vector and the control step behind it are ONE
// launch of one CTA — instead of ncclAllReduce + a control kernel.  The exchange itself is unconditional (every rank
// executes the same sequence of collectives); the control step keeps the early exits of the kernels above.
__device__ __forceinline__ void peer_gave_up(ViCtl *ctl) {
  if (threadIdx.x == 0) { ctl->fault = 2; ctl->done = 1; }
}
__global__ void __launch_bounds__(CTL_NT) peer_em_ctl_kernel(ViArrays a, ViCtl *ctl, int K, int sparse, PeerView pv) {
  __shared__ double scratch[32];
  if (!peer_allreduce_cta<CTL_NT>(a.red, K + RED_EXTRA, pv)) { peer_gave_up(ctl); return; }
  if (ctl->done) return;
  em_ctl_step<CTL_NT>(a, ctl, K, sparse, scratch);
}
__global__ void __launch_bounds__(CTL_NT) peer_rcg_ctl_b_kernel(ViArrays a, ViCtl *ctl, int K, int stage, int stall_on_reject, PeerView pv) {
  __shared__ double scratch[32];
  if (!peer_allreduce_cta<CTL_NT>(a.red, K + 1, pv)) { peer_gave_up(ctl); return; }
  if (ctl->done) return;
  if (stage == 0 ? ctl->stall != 0 : !ctl->didreset) return;
  rcg_ctl_b_step<CTL_NT>(a, ctl, K, stage, stall_on_reject, scratch);
}
__global__ void __launch_bounds__(256) peer_rcgs_ctl_b_kernel(ViArrays va, RcgsGroup g, ViCtl *ctl, int K, int stage, int stall_on_reject, PeerView pv) {
  __shared__ double scratch[32];
  if (!peer_allreduce_cta<256>(va.red, K + 2, pv)) { peer_gave_up(ctl); return; }
  if (ctl->done) return;
  if (stage == 0 ? ctl->stall != 0 : !ctl->didreset) return;
  rcgs_ctl_b_step<256>(va, g, ctl, K, stage, stall_on_reject, scratch);
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
__global__ void dot_kernel(const double *a, const double *b, size_t n, double *block_sums) {
  __shared__ double scratch[32];
  double acc = 0.0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) acc = fma(a[i], b[i], acc);
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

// The same tile from the SPARSE storage, one warp per class.
//   source 0 (RCG): gamma = a_k + b_j off the hits, g_e on them (vi_sparse_rcg.cuh).
//   source 1 (EM) : gamma = logl_jk + dg_k - L_j with logl = l0 off the hits, L_j = M_j + dg_max + log S_j,
//                   S_j = P0_j W + sum_hits dP_e exp(dg_k - dg_max), W = sum_k exp(dg_k - dg_max) (aux[0], aux[1] = dg_max).
__global__ void posterior_tile_sparse_kernel(int source, const uint64_t *__restrict__ nz_ptr, const uint32_t *__restrict__ nz_grp,
                                             const double *__restrict__ hit_val /* sp_g | nz_dP */, const double *__restrict__ nz_logl,
                                             const double *__restrict__ cls_val /* sp_b | P0 */, const double *__restrict__ rowmax,
                                             const double *__restrict__ grp_vec /* a | dg */, const double *__restrict__ aux, double l0,
                                             unsigned long long row0, unsigned long long n_rows, int K, double *__restrict__ out) {
  const int lane = threadIdx.x & 31;
  const unsigned long long warp = (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) >> 5;
  const unsigned long long n_warps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
  for (unsigned long long j = warp; j < n_rows; j += n_warps) {
    const unsigned long long row = row0 + j, a = nz_ptr[row], b = nz_ptr[row + 1];
    if (source == 0) {
      const double bj = cls_val[row];
      for (int k = lane; k < K; k += 32) out[(size_t)k * n_rows + j] = grp_vec[k] + bj;
      __syncwarp();
      for (unsigned long long e = a + lane; e < b; e += 32) out[(size_t)(nz_grp[e] & SP_GRP_MASK) * n_rows + j] = hit_val[e];
    } else {
      const double mx = aux[1];
      double s = 0.0;
      for (unsigned long long e = a + lane; e < b; e += 32) s = fma(hit_val[e], exp(grp_vec[nz_grp[e] & SP_GRP_MASK] - mx), s);
      s = warp_sum(s);
      const double L = rowmax[row] + mx + log(fma(cls_val[row], aux[0], s));
      for (int k = lane; k < K; k += 32) out[(size_t)k * n_rows + j] = l0 + grp_vec[k] - L;
      __syncwarp();
      for (unsigned long long e = a + lane; e < b; e += 32) {
        const int k = (int)(nz_grp[e] & SP_GRP_MASK);
        out[(size_t)k * n_rows + j] = nz_logl[e] + grp_vec[k] - L;
      }
    }
    __syncwarp();
  }
}
// aux[0] = sum_k exp(dg_k - max dg), aux[1] = max dg
__global__ void posterior_aux_kernel(const double *__restrict__ dg, int K, double *__restrict__ aux) {
  __shared__ double scratch[32];
  double mx = -INFINITY;
  for (int k = threadIdx.x; k < K; k += 256) mx = fmax(mx, dg[k]);
  mx = block_max<256>(mx, scratch);
  double w = 0.0;
  for (int k = threadIdx.x; k < K; k += 256) w += exp(dg[k] - mx);
  w = block_sum<256>(w, scratch);
  if (threadIdx.x == 0) { aux[0] = w; aux[1] = mx; }
}

} // namespace mswb

namespace mswb {
// K x n tile of log-posteriors for classes [ec_begin, ec_begin + n) of the shard, on the device (group-major).
void posterior_tile_dev(mswb_ctx *ctx, mswb_lik *lik, uint64_t ec_begin, uint64_t n, double *tile) {
  MSWB_REQUIRE(lik->last_algo >= 0, "no optimisation has been run on this likelihood (posteriors of a bootstrap batch are not kept)");
  const int K = (int)lik->K;
  const int blocks = (int)std::min<uint64_t>(ceil_div(n, 8), (uint64_t)ctx->n_sms * 8);
  if (lik->storage == MSWB_STORE_SPARSE) {
    if (lik->last_algo == MSWB_ALGO_RCG) {
      posterior_tile_sparse_kernel<<<blocks, 256, 0, ctx->stream>>>(0, lik->nz_ptr.p, lik->nz_grp.p, lik->sp_g.p, lik->nz_logl.p, lik->sp_b.p,
                                                                    lik->rowmax.p, lik->last_a.p, nullptr, lik->l0, ec_begin, n, K, tile);
      MSWB_LAUNCHED();
    } else {
      DevBuf<double> aux;
      aux.alloc(2);
      posterior_aux_kernel<<<1, 256, 0, ctx->stream>>>(lik->last_dg.p, K, aux.p);
      MSWB_LAUNCHED();
      posterior_tile_sparse_kernel<<<blocks, 256, 0, ctx->stream>>>(1, lik->nz_ptr.p, lik->nz_grp.p, lik->nz_dP.p, lik->nz_logl.p, lik->P0.p,
                                                                    lik->rowmax.p, lik->last_dg.p, aux.p, lik->l0, ec_begin, n, K, tile);
      MSWB_LAUNCHED();
      MSWB_CUDA(cudaStreamSynchronize(ctx->stream));      // aux goes out of scope
    }
    return;
  }
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
  DevBuf<ViCtl> ctl;
  DevBuf<double> alpha0, N_k, dg, dg_prev, w, red, seg, partials, trace_bound, trace_gnorm, own_counts, block_sums;
  DevBuf<double> cm_const;          // sparse EM: sum_j c_j M_j of this rank's classes (the part of the bound no pass changes)
  DevBuf<unsigned char> trace_reset;
  const double *counts = nullptr;   // device, [N]
  double sum_counts = 0.0;
  std::vector<double> alpha0_host;
  ViArrays arrays{};
  int pstride = 0, grid = 0, max_grid = 0;
  DevBuf<double> rs_a, rs_u, rs_a_new, rs_u_new, rs_ec, rs_mom;   // sparse RCG: group vectors (vi_sparse_rcg.cuh)
  RcgsGroup rs{};
  int tail_max = 16384;          // partial values (grid x columns) the last CTA of a sweep may reduce on its own
  double fx_scale = 1.0;         // fixed-point scale of the sparse pass (vi_kernels.cuh)
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
// Short rows (<= 32 pieces, K <= 64 in fp64 / 128 in fp32): sub-warp row groups.  W lanes x KITER pieces with R rows per
// group and batch; a warp works on (32 / W) x R rows at once and every lane owns one row's division and logarithm.
// MSWB_SMALL=0 falls back to one warp per row (the round-1 shape), for A/B runs.
static bool want_small_shapes() { const char *e = getenv("MSWB_SMALL"); return !(e && e[0] == '0'); }
#define MSWB_TILE_DISPATCH(slots, RMAX, IS_F32, ...)                                         \
  do {                                                                                       \
    const int _s = (int)(slots);                                                             \
    if (_s <= 2 && want_small_shapes()) { using TL = Tile<2, 1, 2>; __VA_ARGS__; }           \
    else if (_s <= 4 && want_small_shapes()) { using TL = Tile<4, 1, 4>; __VA_ARGS__; }      \
    else if (_s <= 8 && want_small_shapes()) { using TL = Tile<8, 1, 8>; __VA_ARGS__; }      \
    else if (_s <= 16 && want_small_shapes()) { using TL = Tile<16, 1, 8>; __VA_ARGS__; }    \
    else if (_s <= 32 && want_small_shapes()) { using TL = Tile<16, 2, 4>; __VA_ARGS__; }    \
    else if (_s <= 32) MSWB_SHAPE_CASE(32, 1, RMAX, __VA_ARGS__)                             \
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
    if (_s <= 2 && want_small_shapes()) { using TL = Tile<2, 1, 2>; __VA_ARGS__; }                          \
    else if (_s <= 4 && want_small_shapes()) { using TL = Tile<4, 1, (RMAX) >= 8 ? 4 : 2>; __VA_ARGS__; }   \
    else if (_s <= 8 && want_small_shapes()) { using TL = Tile<8, 1, (RMAX) >= 8 ? 8 : 4>; __VA_ARGS__; }   \
    else if (_s <= 16 && want_small_shapes()) { using TL = Tile<16, 1, (RMAX) >= 8 ? 8 : 4>; __VA_ARGS__; } \
    else if (_s <= 32 && want_small_shapes()) { using TL = Tile<16, 2, (RMAX) >= 8 ? 4 : 2>; __VA_ARGS__; } \
    else if (_s <= 32) MSWB_SHAPE_CASE(32, 1, RMAX, __VA_ARGS__)                             \
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
  ensure_dyn_smem(kern, ctx->device, smem);
  int per_sm;
  {
    std::lock_guard<std::mutex> lock(mu);
    const Key key{(const void *)kern, nt, smem};
    auto it = cache.find(key);
    if (it == cache.end()) {
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

// How the partial vectors of a sweep with `grid` CTAs and `nvals` columns get summed: by the sweep's last CTA when that
// is a small job (2: and the control step with it, one GPU; 1: several GPUs), else by finalize_ctl_kernel (0).
int tail_mode(const mswb_vi *vi, int grid, int nvals) {
  if ((long long)grid * nvals > vi->tail_max) return 0;
  return vi->ctx->world == 1 ? 2 : 1;
}
// Small problems are bound by launch latency, not by bytes: cap the grid at one CTA per SM there so that the last
// CTA's reduction stays a few microseconds.
int grid_cap(const mswb_vi *vi, int nvals) {
  const mswb_lik *L = vi->lik;
  const bool small = (uint64_t)L->N * L->K * 8 <= ((uint64_t)64 << 20);
  if (!small) return vi->max_grid;
  if (const char *e = getenv("MSWB_SMALL_GRID")) return std::max(1, atoi(e));
  return std::max(1, std::min(vi->max_grid, std::max(vi->ctx->n_sms, vi->tail_max / nvals)));
}

// ctl_mode: -1 none, 0 EM dense, 1 EM sparse, 2 RCG stage 0 (finalize_ctl_kernel)
// peer: the last CTA exchanges the reduced vector over peer memory before its control step (several GPUs)
void launch_finalize(mswb_vi *vi, int nvals, int ctl_mode, int ignore_stall = 0, int peer = 0) {
  finalize_ctl_kernel<<<finalize_grid(nvals, FIN_NT), FIN_NT, 0, vi->ctx->stream>>>(
      vi->partials.p, vi->pstride, vi->grid, nvals, vi->arrays, vi->ctl.p, vi->K, ctl_mode, ignore_stall, peer, vi->ctx->peer);
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

// returns the tail mode the sweep was launched with
template <typename ST, class TL> int launch_em(mswb_vi *vi, const ST *P, int ld) {
  mswb_lik *L = vi->lik;
  cudaStream_t s = vi->ctx->stream;
  const int nvals = vi->K + RED_EXTRA, cap = grid_cap(vi, nvals);
  PipeGeom geom{0, 0, 0};
  if constexpr (TL::TPR == 256) {
    if (want_em_pipe()) geom = pipe_geometry((size_t)ld * sizeof(ST), TL::G * TL::R, 1, TL::TPR);
    if (geom.stages) {
      auto kern = em_lin_pass_kernel<ST, TL, true, true>;
      const size_t smem = pipe_smem_bytes(geom, 1);
      vi->grid = persistent_grid(vi->ctx, kern, TL::NT, smem, ceil_div(L->N_pad, (uint64_t)geom.stage_rows), cap);
      const int tail = tail_mode(vi, vi->grid, nvals);
      kern<<<vi->grid, TL::NT, smem, s>>>(P, ld, L->rowmax.p, vi->counts, vi->arrays, vi->ctl.p, vi->partials.p, vi->pstride, L->N_pad,
                                          vi->K, geom, tail);
      MSWB_LAUNCHED();
      return tail;
    }
  }
  if (em_chunked(L)) geom.stage_rows = 1;
  auto kern = em_lin_pass_kernel<ST, TL, false, false>;      // the hot-loop-only instantiation (see the kernel)
  vi->grid = persistent_grid(vi->ctx, kern, TL::NT, 0, L->N_pad / (TL::G * TL::R), cap);
  const int tail = tail_mode(vi, vi->grid, nvals);
  if (tail == 0) {
    kern<<<vi->grid, TL::NT, 0, s>>>(P, ld, L->rowmax.p, vi->counts, vi->arrays, vi->ctl.p, vi->partials.p, vi->pstride, L->N_pad, vi->K,
                                     geom, 0);
  } else {
    em_lin_pass_kernel<ST, TL, false, true><<<vi->grid, TL::NT, 0, s>>>(P, ld, L->rowmax.p, vi->counts, vi->arrays, vi->ctl.p, vi->partials.p,
                                                                        vi->pstride, L->N_pad, vi->K, geom, tail);
  }
  MSWB_LAUNCHED();
  return tail;
}

template <class TL> void launch_sweep_a(mswb_vi *vi) {
  mswb_lik *L = vi->lik;
  cudaStream_t s = vi->ctx->stream;
  const int ld = (int)L->Kp;
  const int cap = grid_cap(vi, vi->K + 1);
  PipeGeom geom{0, 0, 0};
  if constexpr ((TL::TPR >= 64 || (TL::TPR == 32 && TL::KITER == 2)) && TL::NT <= 512) {
    if (want_rcg_pipe()) geom = pipe_geometry((size_t)ld * 8, TL::G * TL::R, 2, TL::TPR, TL::NT <= 256 ? RCG_RING_BYTES : SMEM_BUDGET, TL::NT <= 256 ? RCG_STAGE_BYTES : 2 * RCG_STAGE_BYTES);
    if (geom.stages) {
      auto kern = rcg_sweep_a_kernel<TL, true>;
      const size_t smem = pipe_smem_bytes(geom, 2);
      const int ga = persistent_grid(vi->ctx, kern, TL::NT, smem, ceil_div(L->N, (uint64_t)geom.stage_rows), cap);
      kern<<<ga, TL::NT, smem, s>>>(L->logl.p, L->gamma.p, ld, vi->arrays, vi->ctl.p, vi->partials.p, vi->pstride, L->N, vi->K, geom);
      MSWB_LAUNCHED();
      return;
    }
  }
  auto kern = rcg_sweep_a_kernel<TL, false>;
  const int ga = persistent_grid(vi->ctx, kern, TL::NT, 0, ceil_div(L->N, (uint64_t)TL::G * TL::R), cap);
  kern<<<ga, TL::NT, 0, s>>>(L->logl.p, L->gamma.p, ld, vi->arrays, vi->ctl.p, vi->partials.p, vi->pstride, L->N, vi->K, geom);
  MSWB_LAUNCHED();
}

// force_tail >= 0 overrides the tail mode (the restart sweep always reduces in its last CTA: it is rare).
// Returns the tail mode used.
template <class TL, int MODE, bool WRITE> int launch_sweep_b(mswb_vi *vi, int only_if_reset, int force_tail = -1) {
  mswb_lik *L = vi->lik;
  cudaStream_t s = vi->ctx->stream;
  const int ld = (int)L->Kp;
  const int nvals = vi->K + 1, cap = grid_cap(vi, nvals);
  double *gam = WRITE || MODE == 0 ? L->gamma.p : nullptr, *stp = MODE == 0 ? L->step.p : nullptr;
  PipeGeom geom{0, 0, 0};
  if constexpr ((TL::TPR >= 64 || (TL::TPR == 32 && TL::KITER == 2)) && TL::NT <= 512) {
    if (want_rcg_pipe()) geom = pipe_geometry((size_t)ld * 8, TL::G * TL::R, 3, TL::TPR, TL::NT <= 256 ? RCG_RING_BYTES : SMEM_BUDGET, TL::NT <= 256 ? RCG_STAGE_BYTES : 2 * RCG_STAGE_BYTES);
    if (geom.stages) {
      auto kern = rcg_sweep_b_kernel<TL, MODE, WRITE, true, MODE == 1>;     // the restart sweep always carries the tail
      const size_t smem = pipe_smem_bytes(geom, 3);
      vi->grid = persistent_grid(vi->ctx, kern, TL::NT, smem, ceil_div(L->N, (uint64_t)geom.stage_rows), cap);
      const int tail = force_tail >= 0 ? force_tail : tail_mode(vi, vi->grid, nvals);
      if (MODE == 1 || tail == 0) {
        kern<<<vi->grid, TL::NT, smem, s>>>(L->logl.p, gam, stp, ld, vi->arrays, vi->counts, vi->ctl.p, vi->partials.p, vi->pstride, L->N,
                                            vi->K, only_if_reset, geom, tail);
      } else {
        auto kern_t = rcg_sweep_b_kernel<TL, MODE, WRITE, true, true>;
        ensure_dyn_smem(kern_t, vi->ctx->device, smem);
        kern_t<<<vi->grid, TL::NT, smem, s>>>(L->logl.p, gam, stp, ld, vi->arrays, vi->counts, vi->ctl.p, vi->partials.p, vi->pstride, L->N,
                                              vi->K, only_if_reset, geom, tail);
      }
      MSWB_LAUNCHED();
      return tail;
    }
  }
  auto kern = rcg_sweep_b_kernel<TL, MODE, WRITE, false, MODE == 1>;
  vi->grid = persistent_grid(vi->ctx, kern, TL::NT, 0, ceil_div(L->N, (uint64_t)TL::G * TL::R), cap);
  const int tail = force_tail >= 0 ? force_tail : tail_mode(vi, vi->grid, nvals);
  if (MODE == 1 || tail == 0) {
    kern<<<vi->grid, TL::NT, 0, s>>>(L->logl.p, gam, stp, ld, vi->arrays, vi->counts, vi->ctl.p, vi->partials.p, vi->pstride, L->N, vi->K,
                                     only_if_reset, geom, tail);
  } else {
    rcg_sweep_b_kernel<TL, MODE, WRITE, false, true><<<vi->grid, TL::NT, 0, s>>>(L->logl.p, gam, stp, ld, vi->arrays, vi->counts, vi->ctl.p,
                                                                                 vi->partials.p, vi->pstride, L->N, vi->K, only_if_reset, geom, tail);
  }
  MSWB_LAUNCHED();
  return tail;
}

// One EM / VB iteration.  One GPU: the pass (its last CTA reduces and takes the control step) — ONE launch — or the
// pass + finalize_ctl_kernel when the partial vectors are too many for one CTA.  Several GPUs: pass, [finalize],
// all-reduce of K + 3 doubles, control kernel.
void em_iteration(mswb_vi *vi) {
  mswb_lik *L = vi->lik;
  mswb_ctx *ctx = vi->ctx;
  cudaStream_t s = ctx->stream;
  const int K = vi->K, nvals = K + RED_EXTRA;
  const int sparse = L->storage == MSWB_STORE_SPARSE ? 1 : 0;
  int tail = 0;
  PassTimer timer(vi);
  if (sparse) {
    const size_t smem = em_sparse_smem_bytes(K);
    MSWB_REQUIRE(smem <= 200 * 1024, "too many groups for the sparse EM pass (weights and accumulators live in shared memory)");
    vi->grid = persistent_grid(ctx, em_sparse_pass_kernel<false>, SP_NT, smem, ceil_div(L->N, (uint64_t)SP_NT), grid_cap(vi, nvals));
    tail = tail_mode(vi, vi->grid, nvals);
    auto kern = tail ? em_sparse_pass_kernel<true> : em_sparse_pass_kernel<false>;
    if (tail) ensure_dyn_smem(kern, ctx->device, smem);
    kern<<<vi->grid, SP_NT, smem, s>>>(L->nz_ptr.p, L->nz_grp.p, L->nz_dP.p, L->P0.p, vi->cm_const.p, vi->counts,
                                      vi->arrays, vi->ctl.p, vi->partials.p, vi->pstride, L->N, L->nnz, K, vi->fx_scale, tail);
    MSWB_LAUNCHED();
  } else if (L->storage == MSWB_STORE_F32) {
    MSWB_TILE_DISPATCH(L->Kp32 / 4, 8, true, tail = launch_em<float, TL>(vi, L->P32.p, (int)L->Kp32));
  } else {
    MSWB_TILE_DISPATCH(L->Kp / 2, 8, false, tail = launch_em<double, TL>(vi, L->P64.p, (int)L->Kp));
  }
  timer.stop();
  if (ctx->world == 1) {
    if (tail == 0) launch_finalize(vi, nvals, sparse);
    return;
  }
  if (ctx->peer_ok) {
    // peer memory: the exchange and the control step ride in the last CTA of the reduction (2 launches per iteration),
    // or in one control CTA behind a sweep whose own last CTA reduced (small problems)
    if (tail == 0) { launch_finalize(vi, nvals, sparse, 0, 1); return; }
    peer_em_ctl_kernel<<<1, CTL_NT, 0, s>>>(vi->arrays, vi->ctl.p, K, sparse, ctx->peer);
    MSWB_LAUNCHED();
    return;
  }
  if (tail == 0) launch_finalize(vi, nvals, -1);
  ctx->allreduce_sum(vi->red.p, nvals);
  em_ctl_kernel<<<1, CTL_NT, 0, s>>>(vi->arrays, vi->ctl.p, K, sparse);
  MSWB_LAUNCHED();
}

// One RCG iteration.
//   one GPU     : sweep A (its last CTA sums the gradient norm), sweep B (computes the Fletcher-Reeves ratio itself;
//                 its last CTA, or finalize_ctl_kernel, reduces and takes the control step), then the restart sweep,
//                 which exits at once unless the control step rejected the move: 3-4 launches, no host round trip.
//   several GPUs: sweep A, all-reduce(1), sweep B, [finalize], all-reduce(K + 1), control kernel: two collectives.
//                 A rejected move sets ctl->stall; everything enqueued behind it exits at once and the host enqueues
//                 the restart (sweep, reduction, all-reduce, control) at its next poll — restarts are rare, and a
//                 third collective per iteration is not.
void rcg_iteration(mswb_vi *vi) {
  mswb_lik *L = vi->lik;
  mswb_ctx *ctx = vi->ctx;
  cudaStream_t s = ctx->stream;
  const int K = vi->K;
  const int slots = L->Kp / 2;
  {
    PassTimer timer(vi);
    MSWB_TILE_DISPATCH_RCG(slots, 8, launch_sweep_a<TL>(vi));   // (one row per batch / three CTAs per SM measured the same: 5456 vs 5445 GB/s)
    timer.stop();
  }
  ctx->allreduce_sum(vi->red.p + K + RED_AUX, 1);
  int tail = 0;
  {
    PassTimer timer(vi);
    MSWB_TILE_DISPATCH_RCG(slots, 4, tail = (launch_sweep_b<TL, 0, true>(vi, 0)));
    timer.stop();
  }
  if (ctx->world == 1) {
    if (tail == 0) launch_finalize(vi, K + 1, 2);
    MSWB_TILE_DISPATCH_RCG(slots, 4, (launch_sweep_b<TL, 1, true>(vi, 1, 2)));
    return;
  }
  if (ctx->peer_ok) {
    if (tail == 0) { launch_finalize(vi, K + 1, 2, 0, 1); return; }
    peer_rcg_ctl_b_kernel<<<1, CTL_NT, 0, s>>>(vi->arrays, vi->ctl.p, K, 0, 1, ctx->peer);
    MSWB_LAUNCHED();
    return;
  }
  if (tail == 0) launch_finalize(vi, K + 1, -1);
  ctx->allreduce_sum(vi->red.p, K + 1);
  rcg_ctl_b_kernel<<<1, CTL_NT, 0, s>>>(vi->arrays, vi->ctl.p, K, 0, 1);
  MSWB_LAUNCHED();
}

// One RCG iteration on the sparse storage (vi_sparse_rcg.cuh): sweep A, [all-reduce(1)], sweep B (the K-sized group part of
// the step in its prologue; its last CTA — or rcgs_finalize_kernel — reduces and, on one GPU, takes the control step), then
// on one GPU the restart sweep, which exits at once unless the step was rejected: 3 launches.
template <int MODE> int launch_rcgs_sweep_b(mswb_vi *vi, int force_tail) {
  mswb_lik *L = vi->lik;
  const int K = vi->K;
  const size_t smem = rcgs_sweep_b_smem(K);
  MSWB_REQUIRE(smem <= SMEM_BUDGET, "too many groups for the sparse RCG sweep (group vectors and accumulators live in shared memory)");
  auto kern0 = rcgs_sweep_b_kernel<MODE, false>;
  auto kern1 = rcgs_sweep_b_kernel<MODE, true>;
  vi->grid = persistent_grid(vi->ctx, kern1, RS_NT, smem, ceil_div(L->N, (uint64_t)RS_NT), grid_cap(vi, K + 2));
  const int tail = force_tail >= 0 ? force_tail : tail_mode(vi, vi->grid, K + 2);
  if (tail == 0) {
    ensure_dyn_smem(kern0, vi->ctx->device, smem);
    kern0<<<vi->grid, RS_NT, smem, vi->ctx->stream>>>(L->nz_ptr.p, L->nz_grp.p, L->nz_logl.p, vi->counts, L->sp_b.p, L->sp_v.p, L->sp_g.p, L->sp_t.p,
                                                       vi->arrays, vi->rs, vi->ctl.p, vi->partials.p, vi->pstride, L->N, L->nnz, K, L->l0, vi->fx_scale, 0);
  } else {
    kern1<<<vi->grid, RS_NT, smem, vi->ctx->stream>>>(L->nz_ptr.p, L->nz_grp.p, L->nz_logl.p, vi->counts, L->sp_b.p, L->sp_v.p, L->sp_g.p, L->sp_t.p,
                                                       vi->arrays, vi->rs, vi->ctl.p, vi->partials.p, vi->pstride, L->N, L->nnz, K, L->l0, vi->fx_scale, tail);
  }
  MSWB_LAUNCHED();
  return tail;
}

void rcgs_iteration(mswb_vi *vi) {
  mswb_lik *L = vi->lik;
  mswb_ctx *ctx = vi->ctx;
  cudaStream_t s = ctx->stream;
  const int K = vi->K;
  {
    PassTimer timer(vi);
    const size_t smem = rcgs_sweep_a_smem(K);
    MSWB_REQUIRE(smem <= SMEM_BUDGET, "too many groups for the sparse RCG sweep");
    const int ga = persistent_grid(ctx, rcgs_sweep_a_kernel, RS_NT, smem, ceil_div(L->N, (uint64_t)RS_NT), grid_cap(vi, K + 2));
    rcgs_sweep_a_kernel<<<ga, RS_NT, smem, s>>>(L->nz_ptr.p, L->nz_grp.p, L->nz_logl.p, L->sp_b.p, L->sp_g.p, vi->arrays, vi->rs, vi->ctl.p,
                                               vi->partials.p, vi->pstride, L->N, L->nnz, K, L->l0);
    MSWB_LAUNCHED();
    timer.stop();
  }
  ctx->allreduce_sum(vi->red.p + K + RS_NORM, 1);
  int tail;
  {
    PassTimer timer(vi);
    tail = launch_rcgs_sweep_b<0>(vi, -1);
    timer.stop();
  }
  if (ctx->world == 1) {
    if (tail == 0) {
      rcgs_finalize_kernel<<<finalize_grid(K + 2, FIN_NT), FIN_NT, 0, s>>>(vi->partials.p, vi->pstride, vi->grid, vi->arrays, vi->rs, vi->ctl.p, K, 0, 0, 0, ctx->peer);
      MSWB_LAUNCHED();
    }
    launch_rcgs_sweep_b<1>(vi, 2);
    return;
  }
  if (tail == 0) {
    const int fused = ctx->peer_ok ? 1 : 0;
    rcgs_finalize_kernel<<<finalize_grid(K + 2, FIN_NT), FIN_NT, 0, s>>>(vi->partials.p, vi->pstride, vi->grid, vi->arrays, vi->rs, vi->ctl.p, K,
                                                                 fused ? 0 : -1, 0, fused, ctx->peer);
    MSWB_LAUNCHED();
    if (fused) return;
  }
  if (ctx->peer_ok) {
    peer_rcgs_ctl_b_kernel<<<1, 256, 0, s>>>(vi->arrays, vi->rs, vi->ctl.p, K, 0, 1, ctx->peer);
    MSWB_LAUNCHED();
    return;
  }
  ctx->allreduce_sum(vi->red.p, K + 2);
  rcgs_ctl_b_kernel<<<1, 256, 0, s>>>(vi->arrays, vi->rs, vi->ctl.p, K, 0, 1);
  MSWB_LAUNCHED();
}

int coop_launch_supported(mswb_ctx *ctx) {
  static std::map<int, int> coop;                    // device -> cooperative launches supported
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  auto it = coop.find(ctx->device);
  if (it == coop.end()) {
    int v = 0;
    MSWB_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrCooperativeLaunch, ctx->device));
    it = coop.emplace(ctx->device, v).first;
  }
  return it->second;
}

// The same for EM / VB on the sparse storage (ems_fused_kernel): passes of up to a few hundred microseconds.
bool ems_fusable(mswb_vi *vi, int *grid_out, size_t *smem_out) {
  mswb_lik *L = vi->lik;
  mswb_ctx *ctx = vi->ctx;
  const int K = vi->K;
  if (vi->opts.algo != MSWB_ALGO_EM || L->storage != MSWB_STORE_SPARSE || ctx->world != 1) return false;
  if (const char *e = getenv("MSWB_FUSED")) if (e[0] == '0') return false;
  if (!coop_launch_supported(ctx)) return false;
  const size_t smem = em_sparse_smem_bytes(K);
  if (smem > 200 * 1024) return false;
  uint64_t max_bytes = (uint64_t)1 << 30;
  if (const char *e = getenv("MSWB_FUSED_MAX_MB")) max_bytes = (uint64_t)atoll(e) << 20;
  if (L->nnz * 12 + L->N * 32 > max_bytes) return false;
  const int grid = persistent_grid(ctx, ems_fused_kernel, SP_NT, smem, ceil_div(L->N, (uint64_t)SP_NT), grid_cap(vi, K + RED_EXTRA));
  if (grid_out) *grid_out = grid;
  if (smem_out) *smem_out = smem;
  return true;
}
bool ems_fused_steps(mswb_vi *vi, uint64_t n) {
  int grid = 0;
  size_t smem = 0;
  if (n == 0 || !ems_fusable(vi, &grid, &smem)) return false;
  mswb_lik *L = vi->lik;
  const int K = vi->K;
  vi->grid = grid;
  const uint64_t *nz_ptr = L->nz_ptr.p; const uint32_t *nz_grp = L->nz_grp.p; const double *nz_dP = L->nz_dP.p;
  const double *P0 = L->P0.p, *rowmax = vi->cm_const.p /* sum_j c_j M_j */, *counts = vi->counts;
  ViArrays va = vi->arrays; ViCtl *ctl = vi->ctl.p;
  double *partials = vi->partials.p; int pstride = vi->pstride;
  unsigned long long N = L->N, nnz = L->nnz, steps = n;
  int Kk = K; double fx = vi->fx_scale;
  int coop_reduce = tail_mode(vi, grid, K + RED_EXTRA) == 2 ? 0 : 1;
  void *args[] = {&nz_ptr, &nz_grp, &nz_dP, &P0, &rowmax, &counts, &va, &ctl, &partials, &pstride, &N, &nnz, &Kk, &fx, &steps, &coop_reduce};
  PassTimer timer(vi);
  MSWB_CUDA(cudaLaunchCooperativeKernel((const void *)ems_fused_kernel, dim3(grid), dim3(SP_NT), args, smem, vi->ctx->stream));
  MSWB_LAUNCHED();
  timer.stop();
  return true;
}

// Small problems on one GPU: up to n iterations in one cooperative launch (rcgs_fused_kernel).  Returns false when the
// problem is not of that kind (the caller then enqueues the iterations one by one).  MSWB_FUSED=0 turns it off.
bool rcgs_fusable(mswb_vi *vi, int *grid_out, size_t *smem_out) {
  mswb_lik *L = vi->lik;
  mswb_ctx *ctx = vi->ctx;
  const int K = vi->K;
  if (vi->opts.algo != MSWB_ALGO_RCG || L->storage != MSWB_STORE_SPARSE || ctx->world != 1) return false;
  if (const char *e = getenv("MSWB_FUSED")) if (e[0] == '0') return false;
  const int can = coop_launch_supported(ctx);
  if (!can) return false;
  const size_t smem = std::max(rcgs_sweep_a_smem(K), rcgs_sweep_b_smem(K));
  if (smem > SMEM_BUDGET) return false;
  // launch latency matters while an iteration's sweeps take up to about a millisecond (64 B per class and per hit, ~2.5 TB/s)
  uint64_t max_bytes = (uint64_t)5 << 29;
  if (const char *e = getenv("MSWB_FUSED_MAX_MB")) max_bytes = (uint64_t)atoll(e) << 20;
  if ((L->nnz + L->N) * 64 > max_bytes) return false;
  const int grid = persistent_grid(ctx, rcgs_fused_kernel, RS_NT, smem, ceil_div(L->N, (uint64_t)RS_NT), grid_cap(vi, K + 2));
  if (grid_out) *grid_out = grid;
  if (smem_out) *smem_out = smem;
  return true;
}
bool rcgs_fused_steps(mswb_vi *vi, uint64_t n) {
  int grid = 0;
  size_t smem = 0;
  if (n == 0 || !rcgs_fusable(vi, &grid, &smem)) return false;
  mswb_lik *L = vi->lik;
  mswb_ctx *ctx = vi->ctx;
  const int K = vi->K;
  vi->grid = grid;
  const uint64_t *nz_ptr = L->nz_ptr.p; const uint32_t *nz_grp = L->nz_grp.p; const double *nz_logl = L->nz_logl.p;
  const double *counts = vi->counts;
  double *sp_b = L->sp_b.p, *sp_v = L->sp_v.p, *sp_g = L->sp_g.p, *sp_t = L->sp_t.p;
  ViArrays va = vi->arrays; RcgsGroup rs = vi->rs; ViCtl *ctl = vi->ctl.p;
  double *partials = vi->partials.p; int pstride = vi->pstride;
  unsigned long long N = L->N, nnz = L->nnz, steps = n;
  int Kk = K; double l0 = L->l0, fx = vi->fx_scale;
  int coop_reduce = tail_mode(vi, grid, K + 2) == 2 ? 0 : 1;      // too many partial vectors for one CTA: every CTA sums its tiles
  void *args[] = {&nz_ptr, &nz_grp, &nz_logl, &counts, &sp_b, &sp_v, &sp_g, &sp_t, &va, &rs, &ctl, &partials, &pstride, &N, &nnz, &Kk, &l0, &fx, &steps,
                  &coop_reduce};
  PassTimer timer(vi);
  MSWB_CUDA(cudaLaunchCooperativeKernel((const void *)rcgs_fused_kernel, dim3(grid), dim3(RS_NT), args, smem, ctx->stream));
  MSWB_LAUNCHED();
  timer.stop();
  return true;
}

void rcgs_restart_stalled(mswb_vi *vi) {
  mswb_ctx *ctx = vi->ctx;
  const int K = vi->K;
  launch_rcgs_sweep_b<1>(vi, 1);
  if (ctx->peer_ok) {
    peer_rcgs_ctl_b_kernel<<<1, 256, 0, ctx->stream>>>(vi->arrays, vi->rs, vi->ctl.p, K, 1, 1, ctx->peer);
    MSWB_LAUNCHED();
    return;
  }
  ctx->allreduce_sum(vi->red.p, K + 2);
  rcgs_ctl_b_kernel<<<1, 256, 0, ctx->stream>>>(vi->arrays, vi->rs, vi->ctl.p, K, 1, 1);
  MSWB_LAUNCHED();
}

// Several GPUs: the restart of a stalled optimisation (every rank stalls at the same iteration: the decision is taken
// on all-reduced values).
void rcg_restart_stalled(mswb_vi *vi) {
  mswb_lik *L = vi->lik;
  mswb_ctx *ctx = vi->ctx;
  const int K = vi->K;
  if (L->storage == MSWB_STORE_SPARSE) { rcgs_restart_stalled(vi); return; }
  const int slots = L->Kp / 2;
  MSWB_TILE_DISPATCH_RCG(slots, 4, (launch_sweep_b<TL, 1, true>(vi, 1, 1)));
  if (ctx->peer_ok) {
    peer_rcg_ctl_b_kernel<<<1, CTL_NT, 0, ctx->stream>>>(vi->arrays, vi->ctl.p, K, 1, 1, ctx->peer);
    MSWB_LAUNCHED();
    return;
  }
  ctx->allreduce_sum(vi->red.p, K + 1);
  rcg_ctl_b_kernel<<<1, CTL_NT, 0, ctx->stream>>>(vi->arrays, vi->ctl.p, K, 1, 1);
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
  for (;;) {
    d2h(&c, vi->ctl.p, 1, vi->ctx->stream);
    MSWB_CUDA(cudaStreamSynchronize(vi->ctx->stream));
    if (!c.stall || c.done) break;
    rcg_restart_stalled(vi);
  }
  if (vi->opts.time_kernels) {
    for (size_t i = 0; i < vi->events_used; ++i) {
      float ms = 0.f;
      MSWB_CUDA(cudaEventElapsedTime(&ms, vi->events[i].first, vi->events[i].second));
      vi->pass_ms_sum += ms;
      vi->pass_launches += 1;
    }
    vi->events_used = 0;
  }
  // (the flag travels in the all-reduced vector: every rank sees it at the same iteration and none is left in a collective)
  MSWB_REQUIRE(c.fault != 2, "a peer rank did not arrive at the all-reduce (it failed, was aborted, or the wait timed out: MSWB_PEER_TIMEOUT_S)");
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

    if (opts->algo == MSWB_ALGO_RCG && lik->storage == MSWB_STORE_SPARSE) {
      // RCG on the sparse storage: separable state off the hits (vi_sparse_rcg.cuh)
      lik_ensure_sparse(lik);
      lik->sp_b.ensure(lik->N); lik->sp_v.ensure(lik->N); lik->sp_g.ensure(lik->nnz); lik->sp_t.ensure(lik->nnz);
      const double g0 = std::log(1.0 / (double)K);
      fill_kernel<<<ctx->n_sms * 4, 256, 0, s>>>(lik->sp_b.p, lik->N, g0); MSWB_LAUNCHED();
      fill_kernel<<<ctx->n_sms * 4, 256, 0, s>>>(lik->sp_g.p, lik->nnz, g0); MSWB_LAUNCHED();
      MSWB_CUDA(cudaMemsetAsync(lik->sp_v.p, 0, lik->sp_v.bytes(), s));
      MSWB_CUDA(cudaMemsetAsync(lik->sp_t.p, 0, lik->sp_t.bytes(), s));
      vi->pass_bytes = lik->nnz * 64 + (uint64_t)lik->N * 64;   // sweep A 20 B per hit + 16 B per class, sweep B 44 + 48
    } else if (opts->algo == MSWB_ALGO_RCG) {
      MSWB_REQUIRE(lik->storage == MSWB_STORE_F64, "RCG needs the fp64 log-likelihood (build the likelihood with MSWB_STORE_F64 or MSWB_STORE_SPARSE)");
      lik_ensure_logl(lik);
      const size_t n = (size_t)lik->N * lik->Kp;
      lik->gamma.ensure(n);
      lik->step.ensure(n);
      fill_kernel<<<ctx->n_sms * 4, 256, 0, s>>>(lik->gamma.p, n, std::log(1.0 / (double)K));
      MSWB_LAUNCHED();
      vi->pass_bytes = (uint64_t)lik->N * K * 56 + (uint64_t)lik->N * 8;   // sweep A 16 B + sweep B 40 B per element
    } else {
      if (lik->storage == MSWB_STORE_SPARSE) {
        lik_ensure_sparse(lik);
        vi->pass_bytes = lik->nnz * 12 + (uint64_t)lik->N * 32;              // hits (group + value); per class two ptr words, P0, c_j (M_j enters through a constant)
      } else {
        lik_ensure_linear(lik);
        const uint64_t bl = lik->storage == MSWB_STORE_F32 ? 4 : 8;
        vi->pass_bytes = (uint64_t)lik->N * K * bl + (uint64_t)lik->N * 16;  // P once, c_j and M_j once
      }
    }

    // per-group vectors
    vi->alpha0.alloc(K); vi->N_k.alloc(K); vi->dg.alloc(K); vi->dg_prev.alloc(K); vi->w.alloc(K); vi->red.alloc(K + RED_EXTRA);
    MSWB_CUDA(cudaMemsetAsync(vi->red.p, 0, (K + RED_EXTRA) * sizeof(double), s));
    vi->alpha0_host.assign(alpha0, alpha0 + K);
    for (int k = 0; k < K; ++k) MSWB_REQUIRE(alpha0[k] > 0.0 && std::isfinite(alpha0[k]), "prior counts must be positive");
    h2d(vi->alpha0.p, alpha0, K, s);
    vi->max_grid = ctx->n_sms * 8;
    vi->pstride = (int)round_up(K + RED_EXTRA, 2);
    vi->partials.alloc((size_t)vi->max_grid * vi->pstride);
    vi->seg.alloc((size_t)RED_SEGS * vi->pstride);
    if (const char *e = getenv("MSWB_TAIL_MAX")) vi->tail_max = atoi(e);
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
      ctx->peer_check();
      vi->counts = vi->own_counts.p;
    } else {
      vi->counts = lik->counts.p;
      vi->sum_counts = lik->sum_counts_total;
    }

    if (opts->algo == MSWB_ALGO_EM && lik->storage == MSWB_STORE_SPARSE) {
      // sum_j c_j M_j: the row maxima enter the bound through this constant only, so the pass does not read them
      const int nb = 296;
      DevBuf<double> parts;
      parts.alloc(nb);
      vi->cm_const.alloc(1);
      dot_kernel<<<nb, 256, 0, s>>>(vi->counts, lik->rowmax.p, lik->N, parts.p);
      MSWB_LAUNCHED();
      sum_blocks_kernel<<<1, 256, 0, s>>>(parts.p, nb, vi->cm_const.p);
      MSWB_LAUNCHED();
      MSWB_CUDA(cudaStreamSynchronize(s));      // (parts goes out of scope)
    }

    // bound constant: lgamma(sum alpha0) - lgamma(sum alpha0 + sum c) - sum lgamma(alpha0)
    long double a_sum = 0.0L, lg_sum = 0.0L;
    for (int k = 0; k < K; ++k) { a_sum += alpha0[k]; lg_sum += std::lgamma(alpha0[k]); }
    const double bconst = (double)(std::lgamma((double)a_sum) - std::lgamma((double)(a_sum + vi->sum_counts)) - lg_sum);

    vi->arrays = ViArrays{vi->alpha0.p, vi->N_k.p, vi->dg.p, vi->w.p, vi->dg_prev.p, vi->red.p, vi->seg.p, vi->trace_bound.p,
                          vi->trace_gnorm.p, vi->trace_reset.p, cap};
    // sparse pass: responsibilities are summed in 64-bit fixed point; |term| <= c_j, so 2^61 / 2^ceil(log2(sum c)) cannot overflow
    vi->fx_scale = std::ldexp(1.0, 61 - (int)std::ceil(std::log2(std::max(1.0, vi->sum_counts))));
    vi_init_kernel<<<1, CTL_NT, 0, s>>>(vi->arrays, vi->ctl.p, K, opts->algo, opts->tol, opts->max_iters, bconst, vi->sum_counts);
    MSWB_LAUNCHED();
    if (opts->algo == MSWB_ALGO_RCG && lik->storage == MSWB_STORE_SPARSE) {
      vi->rs_a.alloc(K); vi->rs_u.alloc(K); vi->rs_a_new.alloc(K); vi->rs_u_new.alloc(K); vi->rs_ec.alloc(K); vi->rs_mom.alloc(RCGS_MOM);
      MSWB_CUDA(cudaMemsetAsync(vi->rs_mom.p, 0, RCGS_MOM * sizeof(double), s));
      vi->rs = RcgsGroup{vi->rs_a.p, vi->rs_u.p, vi->rs_a_new.p, vi->rs_u_new.p, vi->rs_ec.p, vi->rs_mom.p};
      rcgs_init_groups_kernel<<<1, 256, 0, s>>>(vi->arrays, vi->rs, K);
      MSWB_LAUNCHED();
    }
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
    const bool fused = rcgs_fused_steps(vi, n_iters) || ems_fused_steps(vi, n_iters);
    for (uint64_t i = 0; i < n_iters && !fused; ++i) {
      if (vi->opts.algo == MSWB_ALGO_RCG) { if (vi->lik->storage == MSWB_STORE_SPARSE) rcgs_iteration(vi); else rcg_iteration(vi); }
      else em_iteration(vi);
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
    if (vi->opts.algo == MSWB_ALGO_RCG && vi->lik->storage == MSWB_STORE_SPARSE) {
      vi->lik->last_a.ensure(vi->K);
      MSWB_CUDA(cudaMemcpyAsync(vi->lik->last_a.p, vi->rs_a.p, vi->K * sizeof(double), cudaMemcpyDeviceToDevice, vi->ctx->stream));
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
    // (a fused launch stops by itself once the optimiser is done: nothing is wasted by asking for many iterations at a time)
    const uint64_t every = opts->poll_every ? opts->poll_every : ((rcgs_fusable(vi, nullptr, nullptr) || ems_fusable(vi, nullptr, nullptr)) ? 64 : 8);
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
