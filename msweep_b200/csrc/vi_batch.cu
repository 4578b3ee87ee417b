// vi_batch.cu — EM / VB for a BATCH of bootstrap replicates: one sweep of the dense likelihood per iteration serves every
// replicate that is still running (em_lin_batch_kernel, vi_kernels.cuh).  Called from bootstrap.cu for the replicate
// loop of the reference (src/mSWEEP.cpp:496-518) when the algorithm is EM and the storage dense.
#include "handles.cuh"
#include "vi_kernels.cuh"

#include <cmath>
#include <map>
#include <mutex>
#include <tuple>

using namespace mswb;

namespace mswb {
constexpr int CTL_NT = 256;

// One CTA per bootstrap replicate: the same start as vi_init_kernel (EM), on the replicate's own vectors.
__global__ void __launch_bounds__(CTL_NT) vi_init_batch_kernel(ViArrays base, ViCtl *ctls, int K, int pstride, double tol,
                                                               unsigned long long max_iters, double bound_const, double sum_counts) {
  __shared__ double scratch[32];
  const ViArrays a = arrays_of_replicate(base, (int)blockIdx.x, K, pstride);
  ViCtl *ctl = ctls + blockIdx.x;
  double mx = -INFINITY;
  for (int k = threadIdx.x; k < K; k += CTL_NT) {
    const double nk = a.alpha0[k] + sum_counts / (double)K;
    a.N_k[k] = nk;
    const double dg = digamma_series(nk);
    a.dg[k] = dg;
    mx = fmax(mx, dg);
  }
  mx = block_max<CTL_NT>(mx, scratch);
  for (int k = threadIdx.x; k < K; k += CTL_NT) a.w[k] = exp(a.dg[k] - mx);
  if (threadIdx.x == 0) {
    ctl->bound = 0.0; ctl->oldbound = 0.0;
    ctl->oldnorm = 1.0; ctl->newnorm = 0.0; ctl->beta = 0.0;
    ctl->bound_const = bound_const; ctl->tol = tol; ctl->sum_counts = sum_counts; ctl->dg_max = mx;
    ctl->iter = 0; ctl->max_iters = max_iters; ctl->resets = 0;
    ctl->use_old = 0; ctl->didreset = 0; ctl->converged = 0; ctl->fault = 0;
    ctl->stall = 0; ctl->ticket = 0u; ctl->epoch = 0u;
    ctl->done = max_iters == 0 ? 1 : 0;
  }
}

} // namespace mswb

namespace {
constexpr size_t SMEM_BUDGET = 200 * 1024;
// resident CTAs of a kernel on the whole GPU (occupancy query cached per kernel / block size / shared memory)
template <class Kern> int resident_ctas(mswb_ctx *ctx, Kern kern, int nt, size_t smem) {
  struct Key { const void *k; int nt; size_t smem; bool operator<(const Key &o) const { return std::tie(k, nt, smem) < std::tie(o.k, o.nt, o.smem); } };
  static std::map<Key, int> cache;
  static std::mutex mu;
  ensure_dyn_smem(kern, ctx->device, smem);
  std::lock_guard<std::mutex> lock(mu);
  const Key key{(const void *)kern, nt, smem};
  auto it = cache.find(key);
  if (it == cache.end()) {
    int v = 1;
    MSWB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kern, nt, smem));
    it = cache.emplace(key, v < 1 ? 1 : v).first;
  }
  return ctx->n_sms * it->second;
}
} // namespace

// =====================================================================================================
// Batched EM over bootstrap replicates (bootstrap.cu): B count vectors resident on the device, ONE sweep of the
// likelihood per iteration serves every replicate that is still running.
// =====================================================================================================
namespace {

// Replicates per CTA (BT) and rows per batch (R) by row shape: the accumulators (BT x KITER x VEC doubles per thread)
// have to stay in registers.
template <typename ST, int KITER> struct BatchShape {
  static constexpr int VEC = 16 / (int)sizeof(ST);
  static constexpr int BT = KITER * VEC >= 16 ? 2 : 4;
  static constexpr int R = KITER <= 2 ? 4 : 2;     // rows per batch: the reductions and the barrier are paid per batch
};
#define MSWB_BATCH_CASE(ST, TPRV, KITERV, ...) { using TL = Tile<TPRV, KITERV, BatchShape<ST, KITERV>::R>; constexpr int BT = BatchShape<ST, KITERV>::BT; __VA_ARGS__; }
#define MSWB_BATCH_BY_TPR(ST, tpr, KITERV, ...)                                              \
  switch (tpr) {                                                                             \
    case 32: MSWB_BATCH_CASE(ST, 32, KITERV, __VA_ARGS__) break;                             \
    case 64: MSWB_BATCH_CASE(ST, 64, KITERV, __VA_ARGS__) break;                             \
    case 96: MSWB_BATCH_CASE(ST, 96, KITERV, __VA_ARGS__) break;                             \
    case 128: MSWB_BATCH_CASE(ST, 128, KITERV, __VA_ARGS__) break;                           \
    case 160: MSWB_BATCH_CASE(ST, 160, KITERV, __VA_ARGS__) break;                           \
    case 192: MSWB_BATCH_CASE(ST, 192, KITERV, __VA_ARGS__) break;                           \
    case 224: MSWB_BATCH_CASE(ST, 224, KITERV, __VA_ARGS__) break;                           \
    default: MSWB_BATCH_CASE(ST, 256, KITERV, __VA_ARGS__) break;                            \
  }
// rows of up to 1024 pieces (K <= 2048 in fp64, 4096 in fp32); wider rows run the replicates one by one
#define MSWB_BATCH_DISPATCH(ST, slots, ...)                                                  \
  do {                                                                                       \
    const int _s = (int)(slots);                                                             \
    if (_s <= 32) MSWB_BATCH_CASE(ST, 32, 1, __VA_ARGS__)                                    \
    else if (_s <= 512) { const int _t = (int)round_up(ceil_div(_s, 2), 32); MSWB_BATCH_BY_TPR(ST, _t, 2, __VA_ARGS__) }     \
    else { const int _t = (int)round_up(ceil_div(_s, 4), 32); MSWB_BATCH_BY_TPR(ST, _t, 4, __VA_ARGS__) }                    \
  } while (0)

struct BatchRun {
  mswb_ctx *ctx; mswb_lik *lik; int K, pstride, B;
  const double *counts; uint64_t counts_stride;
  ViArrays base; ViCtl *ctls; int *active; double *partials; size_t partials_cap;
  int grid_x = 0;
};

template <typename ST, class TL, int BT> void launch_batch_pass(BatchRun &r, const ST *P, int ld, int n_active) {
  mswb_lik *L = r.lik;
  auto kern = em_lin_batch_kernel<ST, TL, BT>;
  const size_t smem = (size_t)BT * ld * sizeof(ST);
  MSWB_REQUIRE(smem <= SMEM_BUDGET, "too many groups for the batched pass");
  const int n_slices = (n_active + BT - 1) / BT;
  const uint64_t n_batches = L->N_pad / (TL::G * TL::R);
  // all slices resident at once, CTAs of different slices walking the same rows together (the second read hits L2)
  const int resident = resident_ctas(r.ctx, kern, TL::NT, smem);
  int gx = (int)std::max<uint64_t>(1, std::min<uint64_t>(n_batches, (uint64_t)std::max(1, resident / n_slices)));
  MSWB_REQUIRE((size_t)n_slices * BT * gx * r.pstride <= r.partials_cap, "batched pass: partial buffer too small");
  r.grid_x = gx;
  kern<<<dim3(gx, n_slices), TL::NT, smem, r.ctx->stream>>>(P, ld, L->rowmax.p, r.counts, r.counts_stride, r.base, r.active, n_active,
                                                              r.partials, r.pstride, L->N_pad, r.K);
  MSWB_LAUNCHED();
}

} // namespace

// internal (bootstrap.cu).  counts_dev: B x counts_stride doubles (zero beyond the classes), every replicate with the same
// total `sum_counts`.  Returns false when the likelihood's shape has no batched kernel (the caller runs them one by one).
bool mswb_vi_batch_supported(const mswb_lik *lik, const mswb_vi_opts *opts) {
  if (opts->algo != MSWB_ALGO_EM) return false;
  if (const char *e = getenv("MSWB_BOOT_BATCH")) if (e[0] == '0') return false;
  if (lik->storage == MSWB_STORE_F64) return lik->Kp / 2 <= 1024 && lik->ctx->world == 1;
  if (lik->storage == MSWB_STORE_F32) return round_up(lik->K, 4) / 4 <= 1024 && lik->ctx->world == 1;
  return false;
}

int mswb_vi_run_batch_dev_counts(mswb_ctx *ctx, mswb_lik *lik, const double *alpha0, const double *counts_dev,
                                 uint64_t counts_stride, int B, double sum_counts, const mswb_vi_opts *opts, double *thetas,
                                 mswb_vi_stat *stats) {
  return guarded([&] {
    MSWB_REQUIRE(ctx && lik && alpha0 && counts_dev && opts && thetas && B >= 1, "bad arguments");
    MSWB_REQUIRE(mswb_vi_batch_supported(lik, opts), "no batched pass for this likelihood");
    MSWB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    lik_ensure_linear(lik);
    const int K = (int)lik->K, nvals = K + RED_EXTRA;
    for (int k = 0; k < K; ++k) MSWB_REQUIRE(alpha0[k] > 0.0 && std::isfinite(alpha0[k]), "prior counts must be positive");
    BatchRun r{};
    r.ctx = ctx; r.lik = lik; r.K = K; r.B = B; r.counts = counts_dev; r.counts_stride = counts_stride;
    r.pstride = (int)round_up(nvals, 2);
    DevBuf<double> d_alpha0, d_Nk, d_dg, d_dgp, d_w, d_red, d_seg, d_partials;
    DevBuf<ViCtl> d_ctl;
    DevBuf<int> d_active;
    d_alpha0.alloc(K); d_Nk.alloc((size_t)B * K); d_dg.alloc((size_t)B * K); d_dgp.alloc((size_t)B * K); d_w.alloc((size_t)B * K);
    d_red.alloc((size_t)B * nvals); d_seg.alloc((size_t)B * RED_SEGS * r.pstride); d_ctl.alloc(B); d_active.alloc(B);
    r.partials_cap = (size_t)(ctx->n_sms * 8 + 8 * B) * r.pstride;
    d_partials.alloc(r.partials_cap);
    h2d(d_alpha0.p, alpha0, K, s);
    r.base = ViArrays{d_alpha0.p, d_Nk.p, d_dg.p, d_w.p, d_dgp.p, d_red.p, d_seg.p, nullptr, nullptr, nullptr, 0};
    r.ctls = d_ctl.p; r.active = d_active.p; r.partials = d_partials.p;
    long double a_sum = 0.0L, lg_sum = 0.0L;
    for (int k = 0; k < K; ++k) { a_sum += alpha0[k]; lg_sum += std::lgamma(alpha0[k]); }
    const double bconst = (double)(std::lgamma((double)a_sum) - std::lgamma((double)(a_sum + sum_counts)) - lg_sum);
    vi_init_batch_kernel<<<B, CTL_NT, 0, s>>>(r.base, r.ctls, K, r.pstride, opts->tol, opts->max_iters, bconst, sum_counts);
    MSWB_LAUNCHED();

    std::vector<int> active(B);
    for (int b = 0; b < B; ++b) active[b] = b;
    std::vector<ViCtl> ctl_h(B);
    const uint64_t every = opts->poll_every ? opts->poll_every : 16;
    uint64_t passes = 0;
    while (!active.empty() && opts->max_iters > 0) {
      const int n_active = (int)active.size();
      h2d(d_active.p, active.data(), active.size(), s);
      for (uint64_t it = 0; it < every; ++it) {
        if (lik->storage == MSWB_STORE_F32) {
          MSWB_BATCH_DISPATCH(float, lik->Kp32 / 4, (launch_batch_pass<float, TL, BT>(r, lik->P32.p, (int)lik->Kp32, n_active)));
        } else {
          MSWB_BATCH_DISPATCH(double, lik->Kp / 2, (launch_batch_pass<double, TL, BT>(r, lik->P64.p, (int)lik->Kp, n_active)));
        }
        finalize_ctl_batch_kernel<<<dim3(finalize_grid(nvals, 256), n_active), 256, 0, s>>>(r.partials, r.pstride, r.grid_x, nvals, r.base, r.ctls, K,
                                                                                      r.active);
        MSWB_LAUNCHED();
        ++passes;
      }
      d2h(ctl_h.data(), d_ctl.p, B, s);
      MSWB_CUDA(cudaStreamSynchronize(s));     // (also keeps `active` alive until the copy above has been consumed)
      std::vector<int> still;
      for (int b : active) {
        MSWB_REQUIRE(!ctl_h[b].fault, "EM pass: a class normaliser under/overflowed in the linear domain (extreme prior counts)");
        if (!ctl_h[b].done) still.push_back(b);
      }
      active.swap(still);
    }
    std::vector<double> nk((size_t)B * K);
    d2h(nk.data(), d_Nk.p, nk.size(), s);
    d2h(ctl_h.data(), d_ctl.p, B, s);
    MSWB_CUDA(cudaStreamSynchronize(s));
    for (int b = 0; b < B; ++b) {
      for (int k = 0; k < K; ++k) thetas[(size_t)b * K + k] = (nk[(size_t)b * K + k] - alpha0[k]) / sum_counts;
      if (stats) {
        stats[b] = mswb_vi_stat{};
        stats[b].bound = ctl_h[b].bound; stats[b].iters = ctl_h[b].iter; stats[b].converged = ctl_h[b].converged;
        stats[b].pass_launches = passes;
      }
    }
    lik->last_algo = -1;     // the posteriors of a batch are not kept: mswb_vi_posteriors needs a plain run
  });
}

