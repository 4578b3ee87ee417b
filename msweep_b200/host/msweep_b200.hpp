// msweep_b200.hpp — C++17 host-side mirror of the reference's interfaces for the abundance-estimation
// path, implemented on the C ABI of include/msweep_b200.h.  Header-only; this is what mSWEEP's own
// driver would include to gain the B200 backend (see INTEGRATION.md for the exact call sites).
//
// Names follow the reference: Alignment::collapse -> b200::Alignment, LL_WOR21 -> b200::Likelihood,
// rcg_optl() (src/mSWEEP.cpp:176-205) -> b200::rcg_optl(), rcgpar::mixture_components -> the theta the
// optimiser returns, BootstrapSample::resample_counts + the loop at src/mSWEEP.cpp:496-518 ->
// b200::Likelihood::bootstrap().  Errors surface as std::runtime_error carrying mswb_last_error(), so the
// reference's try/catch sites and messages keep working unchanged.
#pragma once
#include "msweep_b200.h"

#include <cstdint>
#include <functional>
#include <limits>
#include <memory>
#include <new>
#include <ostream>
#include <stdexcept>
#include <string>
#include <utility>
#include <cmath>
#include <vector>

namespace b200 {

inline void check(int rc) { if (rc != 0) throw std::runtime_error(mswb_last_error()); }

class Context {
public:
  explicit Context(int device = 0, int rank = 0, int world = 1, const void *nccl_id = nullptr) {
    check(mswb_ctx_create(device, rank, world, nccl_id, nullptr, &h_));
    rank_ = rank; world_ = world;
  }
  // adopt a context created elsewhere (mswb_ctx_create_group)
  Context(mswb_ctx *adopted, int rank, int world) : h_(adopted), rank_(rank), world_(world) {}
  ~Context() { mswb_ctx_destroy(h_); }
  Context(const Context &) = delete;
  Context &operator=(const Context &) = delete;
  mswb_ctx *get() const { return h_; }
  int rank() const { return rank_; }
  int world() const { return world_; }
private:
  mswb_ctx *h_ = nullptr;
  int rank_ = 0, world_ = 1;
};

// std::vector whose resize() leaves trivially constructible elements uninitialised: the parser sizes hundreds of
// megabytes and then fills them from all threads; a zero-fill by one thread first would cost more than the fill.
template <class T> struct default_init_allocator : std::allocator<T> {
  template <class U> struct rebind { using other = default_init_allocator<U>; };
  template <class U, class... A> void construct(U *p, A &&...a) {
    if constexpr (sizeof...(A) == 0) ::new ((void *)p) U; else ::new ((void *)p) U(std::forward<A>(a)...);
  }
};
template <class T> using raw_vector = std::vector<T, default_init_allocator<T>>;

// The strand-merged pseudoalignment as the reference's Alignment holds it after read() (one bit per
// (read, target)), in CSR form: row r = ascending target ids of read r.
struct ReadTable {
  uint64_t n_reads = 0, n_targets = 0;
  std::vector<uint64_t> row_ptr{0};
  raw_vector<uint32_t> targets;
};

// mSWEEP::Alignment after collapse() (include/mSWEEP_alignment.hpp:137-241).
class Alignment {
public:
  Alignment(const Context &ctx, const ReadTable &reads, bool partitioned = false) {
    check((partitioned ? mswb_ec_build_partitioned : mswb_ec_build)(ctx.get(), reads.n_reads, reads.n_targets, reads.row_ptr.data(),
                                                                   reads.targets.data(), &h_));
    check(mswb_ec_info(h_, &n_ecs_, &n_reads_, &n_aligned_, nullptr));
  }
  ~Alignment() { mswb_aln_destroy(h_); }
  Alignment(const Alignment &) = delete;
  Alignment &operator=(const Alignment &) = delete;
  size_t n_ecs() const { return n_ecs_; }          // :217
  size_t n_reads() const { return n_reads_; }      // :218 (all reads, aligned or not)
  size_t n_aligned() const { return n_aligned_; }  // Sample::count_alignments, src/Sample.cpp:52-61
  mswb_aln *get() const { return h_; }
private:
  mswb_aln *h_ = nullptr;
  uint64_t n_ecs_ = 0, n_reads_ = 0, n_aligned_ = 0;
};

struct ViOptions {
  double tol = 1e-6;             // --tol        (src/mSWEEP.cpp:125)
  uint64_t max_iters = 5000;     // --max-iters  (src/mSWEEP.cpp:123)
  int algo = MSWB_ALGO_RCG;
};
struct ViReport { double bound = 0, gnorm = 0; uint64_t iters = 0, resets = 0; bool converged = false; };

// mSWEEP::Likelihood<double> / LL_WOR21 (include/Likelihood.hpp:62-331) plus the optimiser that runs on it.
class Likelihood {
public:
  // ConstructAdaptiveLikelihood (include/Likelihood.hpp:333-380): q, e, min_hits, zero_inflation as there.
  Likelihood(const Context &ctx, const Alignment &aln, const std::vector<uint32_t> &group_of_target,
             const std::vector<uint64_t> &group_sizes, double q, double e, uint64_t min_hits, double zero_inflation,
             int storage = MSWB_STORE_F64) : ctx_(&ctx) {
    check(mswb_lik_build(ctx.get(), aln.get(), group_of_target.data(), (uint32_t)group_sizes.size(), group_sizes.data(), q, e,
                         zero_inflation, min_hits, storage, &h_));
    refresh();
  }
  // Likelihood::from_file (include/Likelihood.hpp:224-252): a precomputed K x N matrix, group-major.
  Likelihood(const Context &ctx, const double *logl, uint32_t n_groups, uint64_t n_ecs, const double *log_counts,
             int storage = MSWB_STORE_F64) : ctx_(&ctx) {
    check(mswb_lik_from_dense(ctx.get(), logl, n_groups, n_ecs, log_counts, storage, &h_));
    refresh();
  }
  ~Likelihood() { mswb_lik_destroy(h_); }
  Likelihood(const Likelihood &) = delete;
  Likelihood &operator=(const Likelihood &) = delete;

  size_t get_rows() const { return n_groups_; }                 // log_mat().get_rows()
  size_t get_cols() const { return n_ecs_; }
  size_t n_groups_all() const { return n_groups_all_; }
  uint64_t ec_begin() const { return ec_begin_; }               // first global class index of this rank's shard
  std::vector<bool> groups_considered() const {                 // include/Likelihood.hpp:79, 330
    std::vector<uint8_t> m(n_groups_all_);
    check(mswb_lik_mask(h_, m.data(), nullptr));
    return std::vector<bool>(m.begin(), m.end());
  }
  mswb_lik *get() const { return h_; }

  // Posteriors on demand (what rcg_optl returns in the reference): K x (end-begin), group-major, log scale.
  std::vector<double> posteriors(uint64_t begin, uint64_t end) const {
    std::vector<double> out((size_t)n_groups_ * (end - begin));
    check(mswb_vi_posteriors(ctx_->get(), h_, begin, end, out.data()));
    return out;
  }

  // Read bins, the job of mGEMS::BinFromMatrix at src/mSWEEP.cpp:437-469: bins[k] = ascending ids of the reads whose
  // class has log-posterior >= log_threshold[k] for kept group k (+inf skips a group).  Classes of this rank's shard.
  std::vector<std::vector<uint32_t>> assign(const Alignment &aln, const std::vector<double> &log_threshold) {
    if (log_threshold.size() != n_groups_) throw std::runtime_error("thresholds must have one value per group");
    std::vector<uint64_t> ptr(n_groups_ + 1);
    check(mswb_vi_assign(ctx_->get(), h_, aln.get(), log_threshold.data(), ptr.data()));
    std::vector<uint32_t> flat(ptr[n_groups_]);
    check(mswb_vi_assign_fetch(h_, flat.data()));
    std::vector<std::vector<uint32_t>> bins(n_groups_);
    for (uint32_t k = 0; k < n_groups_; ++k) bins[k].assign(flat.begin() + ptr[k], flat.begin() + ptr[k + 1]);
    return bins;
  }

  // resample + re-estimate, src/mSWEEP.cpp:496-518; rows of replicates owned by other ranks stay NaN.
  std::vector<std::vector<double>> bootstrap(const std::vector<double> &alpha0, const ViOptions &o, uint64_t iters,
                                             uint64_t bootstrap_count, int32_t seed, int replica_rank = 0,
                                             int replica_world = 1, int rng_mode = MSWB_RNG_LIBSTDCXX_EXACT) {
    std::vector<double> flat((size_t)iters * n_groups_, std::numeric_limits<double>::quiet_NaN());
    mswb_vi_opts opts{o.tol, o.max_iters, o.algo, 0, 0};
    check(mswb_bootstrap_run(ctx_->get(), h_, alpha0.data(), &opts, iters, bootstrap_count, seed, rng_mode, replica_rank,
                             replica_world, flat.data(), nullptr));
    std::vector<std::vector<double>> out(iters);
    for (uint64_t r = 0; r < iters; ++r) out[r].assign(flat.begin() + r * n_groups_, flat.begin() + (r + 1) * n_groups_);
    return out;
  }
private:
  void refresh() { check(mswb_lik_info(h_, &n_groups_all_, &n_groups_, &n_ecs_, &ec_begin_, nullptr)); }
  const Context *ctx_;
  mswb_lik *h_ = nullptr;
  uint32_t n_groups_all_ = 0, n_groups_ = 0;
  uint64_t n_ecs_ = 0, ec_begin_ = 0;
};

// Drop-in for `rcg_optl(args, ll_mat, log_ec_counts, prior_counts, log)` + `rcgpar::mixture_components`
// (src/mSWEEP.cpp:176-205, 419-423): runs the chosen optimiser on the device-resident likelihood and
// returns the relative abundances.  log_ec_counts may be null (use the class counts the likelihood
// holds) or the resampled log-counts of a bootstrap replicate (-inf allowed).  Progress lines go to
// `log` in rcgpar's format every 5th iteration when it is non-null.
inline std::vector<double> rcg_optl(const Context &ctx, Likelihood &ll, const std::vector<double> *log_ec_counts,
                                    const std::vector<double> &prior_counts, const ViOptions &o, std::ostream *log = nullptr,
                                    ViReport *report = nullptr) {
  if (prior_counts.size() != ll.get_rows()) throw std::runtime_error("prior counts must have one value per group");
  mswb_vi_opts opts{o.tol, o.max_iters, o.algo, 0, 0};
  mswb_vi_stat st{};
  std::vector<double> theta(ll.get_rows());
  auto cb = [](void *user, uint64_t iter, double bound, double gnorm) {
    std::ostream &os = *static_cast<std::ostream *>(user);
    if (iter % 5 == 0) os << "  iter: " << iter << ", bound: " << bound << ", |g|: " << gnorm << '\n';
  };
  check(mswb_vi_run(ctx.get(), ll.get(), prior_counts.data(), log_ec_counts ? log_ec_counts->data() : nullptr, &opts,
                    theta.data(), &st, log ? +cb : nullptr, log));
  if (log) *log << std::endl;
  if (report) *report = ViReport{st.bound, st.gnorm, st.iters, st.resets, st.converged != 0};
  return theta;
}

// Signature-compatible form of the rcgpar entry points the reference calls at src/mSWEEP.cpp:194-202:
//   seamat::DenseMatrix<double> rcgpar::rcg_optl_omp(const seamat::Matrix<double>& logl, const std::vector<double>&
//       log_times_observed, const std::vector<double>& alpha0, const double& tol, uint16_t max_iters, std::ostream& log)
// with the matrix types flattened to what they hold: `logl` is the K x N group-major log-likelihood
// (seamat::DenseMatrix row-major storage), the return value the K x N matrix of LOG-posteriors (gamma_Z) the reference
// moves into Sample::ec_probabilities (include/Sample.hpp:66).  algo: MSWB_ALGO_RCG for rcg_optl_omp / rcg_optl_torch,
// MSWB_ALGO_EM for em_torch (storage MSWB_STORE_F32 = its "float" precision).  For matrices a host can hold; the
// device-resident path above is the one that scales.  mixture_components(result, log_times_observed) is `theta_out`.
inline std::vector<double> rcg_optl_dense(const Context &ctx, const std::vector<double> &logl, uint32_t n_groups, uint64_t n_ecs,
                                          const std::vector<double> &log_times_observed, const std::vector<double> &alpha0,
                                          double tol, uint64_t max_iters, std::ostream &log, int algo = MSWB_ALGO_RCG,
                                          int storage = MSWB_STORE_F64, std::vector<double> *theta_out = nullptr) {
  if (logl.size() != (size_t)n_groups * n_ecs) throw std::runtime_error("logl must hold n_groups x n_ecs values");
  if (log_times_observed.size() != n_ecs) throw std::runtime_error("log_times_observed must have one value per equivalence class");
  Likelihood ll(ctx, logl.data(), n_groups, n_ecs, log_times_observed.data(), storage);
  ViOptions o;
  o.tol = tol; o.max_iters = max_iters; o.algo = algo;
  std::vector<double> theta = rcg_optl(ctx, ll, nullptr, alpha0, o, log.good() ? &log : nullptr);   // a never-opened ofstream = quiet (src/mSWEEP.cpp:190)
  if (theta_out) *theta_out = std::move(theta);
  return ll.posteriors(0, n_ecs);
}

// The digamma series mSWEEP carries for RATE (src/Sample.cpp:87-97): recurrence up to 7, then the expansion in 1/(x - 1/2).
inline double digamma(double x) {
  double r = 0.0;
  for (; x < 7.0; x += 1.0) r -= 1.0 / x;
  x -= 0.5;
  const double i2 = 1.0 / (x * x), i4 = i2 * i2;
  return r + std::log(x) + i2 / 24.0 - 7.0 / 960.0 * i4 + 31.0 / 8064.0 * i4 * i2 - 127.0 / 30720.0 * i4 * i4;
}

// --run-rate (Sample::dirichlet_kld + Sample::get_rates, src/Sample.cpp:99-151).  The reference sums exp(gamma) once
// per read of every class, i.e. alphas[k] = the expected read count of group k = theta[k] * counts_total, which
// is what the optimiser already holds; the K x N matrix is not needed.  Returns {log_KLD, RATE}.
inline std::pair<std::vector<double>, std::vector<double>> dirichlet_kld(const std::vector<double> &theta, double counts_total) {
  const size_t K = theta.size();
  std::vector<double> alphas(K), log_kld(K), rate(K);
  double alpha0 = 0.0;
  for (size_t k = 0; k < K; ++k) { alphas[k] = theta[k] * counts_total; alpha0 += alphas[k]; }
  for (size_t k = 0; k < K; ++k) {
    const double a = alphas[k];
    const double kld = std::lgamma(alpha0) - std::lgamma(alpha0 - a) - std::lgamma(a) + a * (digamma(a) - digamma(alpha0));
    log_kld[k] = std::log(std::max(kld, 1e-16));
  }
  double mx = 0.0;                                       // starts at 0, not at the first element (:138-142)
  for (double v : log_kld) mx = mx > v ? mx : v;
  double sum = 0.0;
  for (double v : log_kld) sum += std::exp(v - mx);
  const double lse = std::log(sum) + mx;
  for (size_t k = 0; k < K; ++k) rate[k] = std::exp(log_kld[k] - lse);
  return {log_kld, rate};
}

} // namespace b200
