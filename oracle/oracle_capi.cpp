// oracle/oracle_capi.cpp — flat C entry points over oracle.hpp so that tests/ and bench.py can
// drive the oracle through ctypes.  TEST INFRASTRUCTURE ONLY (see oracle.hpp).
#include "oracle.hpp"

#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <cmath>

#ifdef _OPENMP
#include <omp.h>
#endif

using namespace oracle;

namespace {
thread_local std::string g_err;
template <typename F> int guarded(F &&f) {
  try { f(); return 0; } catch (const std::exception &e) { g_err = e.what(); return 1; }
}
template <typename T> void copy_out(const std::vector<T> &v, T *dst) {
  if (dst && !v.empty()) std::memcpy(dst, v.data(), v.size() * sizeof(T));
}
struct ViOut { ViResult r; };
} // namespace

extern "C" {

const char *orc_last_error() { return g_err.c_str(); }
int orc_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}

// ---- scalar known-answer helpers --------------------------------------------------------------
uint64_t orc_pattern_hash(const uint32_t *targets, uint64_t n) { return pattern_hash(targets, n); }
double orc_digamma(double x) { return digamma_series(x); }
double orc_ldbb_scaled(uint64_t k, uint64_t n, double a, double b) { return ldbb_scaled(k, n, a, b); }
void orc_bb_parameters(uint64_t n, double q, double e, double *a, double *b) { bb_parameters(n, q, e, a, b); }

// ---- read table -------------------------------------------------------------------------------
int orc_reads_from_files(const char *const *paths, int n_paths, uint64_t n_targets, const char *merge_mode, void **out) {
  return guarded([&] {
    std::vector<std::ifstream> files;
    std::vector<std::istream*> strands;
    files.reserve(n_paths);
    for (int i = 0; i < n_paths; ++i) {
      files.emplace_back(paths[i]);
      if (!files.back().good()) throw std::runtime_error(std::string("cannot open ") + paths[i]);
    }
    for (auto &f : files) strands.push_back(&f);
    *out = new ReadTable(read_themisto(strands, n_targets, merge_mode));
  });
}
int orc_reads_from_csr(uint64_t R, uint64_t T, const uint64_t *row_ptr, const uint32_t *targets, void **out) {
  return guarded([&] { *out = new ReadTable(from_csr(R, T, row_ptr, targets)); });
}
uint64_t orc_reads_n(const void *h) { return ((const ReadTable*)h)->n_reads; }
uint64_t orc_reads_nnz(const void *h) {
  uint64_t n = 0;
  for (auto &r : ((const ReadTable*)h)->rows) n += r.size();
  return n;
}
void orc_reads_csr(const void *h, uint64_t *row_ptr, uint32_t *targets) {
  const ReadTable &t = *(const ReadTable*)h;
  uint64_t p = 0;
  for (uint64_t i = 0; i < t.n_reads; ++i) {
    row_ptr[i] = p;
    for (uint32_t x : t.rows[i]) targets[p++] = x;
  }
  row_ptr[t.n_reads] = p;
}
void orc_reads_free(void *h) { delete (ReadTable*)h; }

// ---- EC table ---------------------------------------------------------------------------------
int orc_ec_build(const void *reads, void **out) {
  return guarded([&] { *out = new EcTable(collapse(*(const ReadTable*)reads)); });
}
uint64_t orc_ec_n(const void *h) { return ((const EcTable*)h)->n_ecs(); }
uint64_t orc_ec_pat_nnz(const void *h) { return ((const EcTable*)h)->pat_targets.size(); }
uint64_t orc_ec_n_read_ids(const void *h) { return ((const EcTable*)h)->read_ids.size(); }
void orc_ec_export(const void *h, uint64_t *hash, uint64_t *count, uint32_t *rep_read,
                   uint64_t *pat_ptr, uint32_t *pat_targets, uint64_t *read_ptr, uint32_t *read_ids) {
  const EcTable &e = *(const EcTable*)h;
  copy_out(e.hash, hash); copy_out(e.count, count); copy_out(e.rep_read, rep_read);
  copy_out(e.pat_ptr, pat_ptr); copy_out(e.pat_targets, pat_targets);
  copy_out(e.read_ptr, read_ptr); copy_out(e.read_ids, read_ids);
}
void orc_ec_free(void *h) { delete (EcTable*)h; }

// ---- grouping ---------------------------------------------------------------------------------
int orc_grouping_read(const char *path, void **out) {
  return guarded([&] {
    std::ifstream f(path);
    *out = new Grouping(read_grouping(f));
  });
}
uint32_t orc_grouping_n_groups(const void *h) { return ((const Grouping*)h)->n_groups(); }
uint64_t orc_grouping_n_targets(const void *h) { return ((const Grouping*)h)->group_of_target.size(); }
void orc_grouping_export(const void *h, uint32_t *group_of_target, uint64_t *sizes) {
  const Grouping &g = *(const Grouping*)h;
  copy_out(g.group_of_target, group_of_target); copy_out(g.sizes, sizes);
}
const char *orc_grouping_name(const void *h, uint32_t g) { return ((const Grouping*)h)->names[g].c_str(); }
void orc_grouping_free(void *h) { delete (Grouping*)h; }

// ---- likelihood -------------------------------------------------------------------------------
int orc_lik_build(const void *ec, const uint32_t *group_of_target, uint64_t T, uint32_t K, const uint64_t *sizes,
                  double q, double e, double zero_inflation, uint64_t min_hits, int keep_hit_counts, void **out) {
  return guarded([&] {
    Grouping g;
    g.group_of_target.assign(group_of_target, group_of_target + T);
    g.sizes.assign(sizes, sizes + K);
    g.names.resize(K);
    *out = new Likelihood(build_likelihood(*(const EcTable*)ec, g, q, e, zero_inflation, min_hits, keep_hit_counts != 0));
  });
}
uint32_t orc_lik_n_groups(const void *h) { return ((const Likelihood*)h)->n_groups; }
uint64_t orc_lik_n_ecs(const void *h) { return ((const Likelihood*)h)->n_ecs; }
uint64_t orc_lik_lut_cols(const void *h) { return ((const Likelihood*)h)->lut_cols; }
void orc_lik_export(const void *h, uint8_t *mask, uint64_t *hits, double *logl, double *log_counts,
                    double *lut, uint32_t *hit_counts) {
  const Likelihood &L = *(const Likelihood*)h;
  copy_out(L.groups_mask, mask); copy_out(L.group_hits, hits); copy_out(L.logl, logl);
  copy_out(L.log_counts, log_counts); copy_out(L.lut, lut); copy_out(L.hit_counts, hit_counts);
}
const double *orc_lik_logl_ptr(const void *h) { return ((const Likelihood*)h)->logl.data(); }
const double *orc_lik_log_counts_ptr(const void *h) { return ((const Likelihood*)h)->log_counts.data(); }
void orc_lik_free(void *h) { delete (Likelihood*)h; }

// ---- optimiser --------------------------------------------------------------------------------
// algo: 0 = RCG (rcg_optl_omp), 1 = EM (em_torch double).  logl is K x N group-major.
// stats: [0]=bound, [1]=iters, [2]=converged.  trace_* may be NULL, else capacity max_iters.
int orc_vi_run(int algo, const double *logl, uint32_t K, uint64_t N, const double *log_counts, const double *alpha0,
               double tol, uint64_t max_iters, double *theta, double *N_k, double *gamma, double *stats,
               double *trace_bound, double *trace_gnorm, uint8_t *trace_reset, double *trace_t_end) {
  return guarded([&] {
    ViResult r = algo == 0 ? rcg_optl(logl, K, N, log_counts, alpha0, tol, max_iters)
                           : em_optl(logl, K, N, log_counts, alpha0, tol, max_iters);
    if (theta) copy_out(mixture_components(r.gamma.data(), K, N, log_counts), theta);
    copy_out(r.N_k, N_k);
    copy_out(r.gamma, gamma);
    if (stats) { stats[0] = r.bound; stats[1] = (double)r.iters; stats[2] = r.converged ? 1.0 : 0.0; }
    copy_out(r.trace.bound, trace_bound); copy_out(r.trace.gnorm, trace_gnorm); copy_out(r.trace.reset, trace_reset); copy_out(r.trace.t_end, trace_t_end);
  });
}

// ---- RATE and read bins -----------------------------------------------------------------------
int orc_dirichlet_kld(const double *gamma, uint32_t K, uint64_t N, const double *log_counts, double *log_kld, double *rate) {
  return guarded([&] {
    const RateResult r = dirichlet_kld(gamma, K, N, log_counts);
    copy_out(r.log_kld, log_kld);
    copy_out(r.rate, rate);
  });
}
// bin_ptr: K+1 offsets; read_ids: capacity entries (call with read_ids == NULL first to size it)
int orc_bin_reads(const double *gamma, uint32_t K, uint64_t N, const double *theta, const uint8_t *want, const uint64_t *read_ptr,
                  const uint32_t *read_ids_in, uint64_t *bin_ptr, uint32_t *read_ids) {
  return guarded([&] {
    const auto bins = bin_reads(gamma, K, N, std::vector<double>(theta, theta + K), std::vector<uint8_t>(want, want + K),
                                std::vector<uint64_t>(read_ptr, read_ptr + N + 1), std::vector<uint32_t>(read_ids_in, read_ids_in + read_ptr[N]));
    bin_ptr[0] = 0;
    for (uint32_t k = 0; k < K; ++k) {
      if (read_ids) std::memcpy(read_ids + bin_ptr[k], bins[k].data(), bins[k].size() * sizeof(uint32_t));
      bin_ptr[k + 1] = bin_ptr[k] + bins[k].size();
    }
  });
}

// ---- bootstrap --------------------------------------------------------------------------------
// Sequential replicates from ONE generator, as the reference's loop does (src/mSWEEP.cpp:498-502).
int orc_bootstrap_resample(const uint64_t *ec_counts, uint64_t N, int32_t seed, uint64_t bootstrap_count,
                           uint64_t n_replicates, uint32_t *out /* n_replicates x N */) {
  return guarded([&] {
    Bootstrapper b(std::vector<uint64_t>(ec_counts, ec_counts + N), seed, bootstrap_count);
    for (uint64_t r = 0; r < n_replicates; ++r) {
      const std::vector<uint32_t> c = b.resample_raw();
      std::memcpy(out + r * N, c.data(), N * sizeof(uint32_t));
    }
  });
}

// ---- output -----------------------------------------------------------------------------------
int orc_write_abundances(const char *path, const char *version, uint64_t n_reads, uint64_t n_aligned,
                         const char *const *est_names, uint32_t n_est, const char *const *zero_names, uint32_t n_zero,
                         const double *results /* (iters+1) x n_est */, uint64_t bootstrap_iters) {
  return guarded([&] {
    std::ofstream of(path);
    std::vector<std::string> en(est_names, est_names + n_est), zn;
    if (n_zero) zn.assign(zero_names, zero_names + n_zero);
    std::vector<std::vector<double>> res(bootstrap_iters + 1);
    for (uint64_t b = 0; b <= bootstrap_iters; ++b) res[b].assign(results + b * n_est, results + (b + 1) * n_est);
    write_abundances(of, version, n_reads, n_aligned, en, zn, res, bootstrap_iters);
  });
}

} // extern "C"
