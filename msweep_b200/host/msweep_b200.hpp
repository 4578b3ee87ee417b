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
#include <ostream>
#include <stdexcept>
#include <string>
#include <utility>
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

// The strand-merged pseudoalignment as the reference's Alignment holds it after read() (one bit per
// (read, target)), in CSR form: row r = ascending target ids of read r.
struct ReadTable {
  uint64_t n_reads = 0, n_targets = 0;
  std::vector<uint64_t> row_ptr{0};
  std::vector<uint32_t> targets;
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

} // namespace b200
