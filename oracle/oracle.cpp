// oracle/oracle.cpp — CPU restatement of mSWEEP's abundance-estimation hot path.
// TEST INFRASTRUCTURE ONLY (see oracle.hpp).  Citations are relative to /root/reference/.
#include "oracle.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <limits>
#include <map>
#include <numeric>
#include <ostream>
#include <random>
#include <sstream>
#include <stdexcept>
#include <unordered_map>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace oracle {

// =============================================================================================
// Pseudoalignment parsing
// =============================================================================================

// include/mSWEEP_alignment.hpp:54-94.  Every line is "<read_id> <t0> <t1> ...", single-space
// separated, all 0-based.  The reference sets bit read_id*n_targets + t with no range check on
// t, so a target id >= n_targets lands in a later read's row; the flat index keeps that.
uint64_t parse_plaintext_strand(std::istream &in, uint64_t n_targets, std::vector<uint64_t> *bits) {
  std::string line, tok;
  uint64_t n_lines = 0;
  while (std::getline(in, line)) {
    ++n_lines;
    try {
      std::stringstream ss(line);
      std::getline(ss, tok, ' ');
      const uint64_t read_id = std::stoul(tok);
      while (std::getline(ss, tok, ' ')) bits->push_back(read_id * n_targets + std::stoul(tok));
    } catch (const std::exception &ex) {
      const std::string what(ex.what());
      if (what.find("stoul") != std::string::npos)
        throw std::runtime_error("File format not supported on line " + std::to_string(n_lines) + " with content: " + line);
      throw std::runtime_error("Could not parse line " + std::to_string(n_lines) + " with content: " + line);
    }
  }
  return n_lines;
}

static void sort_unique(std::vector<uint64_t> &v) {
  std::sort(v.begin(), v.end());
  v.erase(std::unique(v.begin(), v.end()), v.end());
}

// include/mSWEEP_alignment.hpp:97-135.
ReadTable read_themisto(const std::vector<std::istream*> &strands, uint64_t n_targets,
                        const std::string &merge_mode) {
  ReadTable out;
  out.n_targets = n_targets;
  std::vector<uint64_t> merged;
  for (size_t s = 0; s < strands.size(); ++s) {
    std::vector<uint64_t> bits;
    out.n_reads = parse_plaintext_strand(*strands[s], n_targets, &bits);   // :121, last strand wins
    sort_unique(bits);
    if (s == 0) {
      merged.swap(bits);
    } else {
      std::vector<uint64_t> tmp;
      if (merge_mode == "intersection")
        std::set_intersection(merged.begin(), merged.end(), bits.begin(), bits.end(), std::back_inserter(tmp));
      else if (merge_mode == "union")
        std::set_union(merged.begin(), merged.end(), bits.begin(), bits.end(), std::back_inserter(tmp));
      else
        throw std::runtime_error("Unrecognized option `" + merge_mode + "` for --themisto-mode");
      merged.swap(tmp);
    }
  }
  out.rows.assign(out.n_reads, {});
  for (uint64_t b : merged) {
    const uint64_t r = b / n_targets;
    if (r < out.n_reads) out.rows[r].push_back((uint32_t)(b % n_targets));   // collapse() only visits i < n_queries
  }
  return out;
}

ReadTable from_csr(uint64_t n_reads, uint64_t n_targets, const uint64_t *row_ptr, const uint32_t *targets) {
  ReadTable out;
  out.n_reads = n_reads;
  out.n_targets = n_targets;
  out.rows.resize(n_reads);
  for (uint64_t i = 0; i < n_reads; ++i) {
    out.rows[i].assign(targets + row_ptr[i], targets + row_ptr[i + 1]);
    std::sort(out.rows[i].begin(), out.rows[i].end());
    out.rows[i].erase(std::unique(out.rows[i].begin(), out.rows[i].end()), out.rows[i].end());
  }
  return out;
}

// =============================================================================================
// Equivalence classes
// =============================================================================================

// include/mSWEEP_alignment.hpp:150-155: order-dependent 64-bit fold over ascending set bits.
uint64_t pattern_hash(const uint32_t *targets, size_t n) {
  uint64_t h = 0;
  for (size_t a = 0; a < n; ++a) h ^= (uint64_t)targets[a] + 0x517cc1b727220a95ULL + (h << 6) + (h >> 2);
  return h;
}

// include/mSWEEP_alignment.hpp:137-215.  Classes are keyed on the HASH (collisions merge, :157-168),
// ordered by ascending hash (std::map, :172-184, :200), read ids ascending inside a class, pattern
// of a class = row of its first read (:204-206).  Reads with no set bit are skipped (:149).
EcTable collapse(const ReadTable &reads) {
  std::map<uint64_t, std::vector<uint32_t>> classes;
  for (uint64_t i = 0; i < reads.n_reads; ++i) {
    const auto &row = reads.rows[i];
    if (row.empty()) continue;
    classes[pattern_hash(row.data(), row.size())].push_back((uint32_t)i);
  }
  EcTable ec;
  ec.n_reads = reads.n_reads;
  ec.n_targets = reads.n_targets;
  ec.read_ptr.push_back(0);
  ec.pat_ptr.push_back(0);
  for (auto &kv : classes) {
    ec.hash.push_back(kv.first);
    ec.count.push_back(kv.second.size());
    ec.rep_read.push_back(kv.second.front());
    ec.read_ids.insert(ec.read_ids.end(), kv.second.begin(), kv.second.end());
    ec.read_ptr.push_back(ec.read_ids.size());
    const auto &row = reads.rows[kv.second.front()];
    ec.pat_targets.insert(ec.pat_targets.end(), row.begin(), row.end());
    ec.pat_ptr.push_back(ec.pat_targets.size());
  }
  return ec;
}

// =============================================================================================
// Grouping
// =============================================================================================

// include/Grouping.hpp:62-83 (ids in order of first appearance, sizes counted per name) and
// include/Reference.hpp:67-94 (columns split on the delimiter, one indicator per line).
Grouping read_grouping(std::istream &in, char delimiter, size_t column) {
  if (!in.good()) throw std::runtime_error("Could not read cluster indicators.");
  Grouping g;
  std::unordered_map<std::string, uint32_t> ids;
  std::string line, field;
  while (std::getline(in, line)) {
    std::stringstream ss(line);
    size_t col = 0;
    bool found = false;
    while (std::getline(ss, field, delimiter)) {
      if (col == column) { found = true; break; }
      ++col;
    }
    if (!found) continue;   // reference: a line with fewer columns contributes nothing to that grouping
    auto it = ids.find(field);
    if (it == ids.end()) {
      it = ids.emplace(field, (uint32_t)g.names.size()).first;
      g.names.push_back(field);
      g.sizes.push_back(0);
    }
    g.sizes[it->second] += 1;
    g.group_of_target.push_back(it->second);
  }
  if (g.group_of_target.empty()) throw std::runtime_error("The grouping contains 0 reference sequences");
  return g;
}

// =============================================================================================
// Likelihood
// =============================================================================================

double lbeta(double x, double y) { return std::lgamma(x) + std::lgamma(y) - std::lgamma(x + y); }   // Likelihood.hpp:47-50

// Likelihood.hpp:52-60: log C(n,k) + lbeta(k+a, n-k+b) - lbeta(n+a, b)   ("scaled": note the normaliser)
double ldbb_scaled(uint64_t k, uint64_t n, double alpha, double beta) {
  const double lbc = std::lgamma((double)n + 1.0) - std::lgamma((double)k + 1.0) - std::lgamma((double)(n - k) + 1.0);
  return lbc + lbeta((double)k + alpha, (double)(n - k) + beta) - lbeta((double)n + alpha, beta);
}

// Likelihood.hpp:198-207 with bb_constants = {q, e} (ctor arg order, :212-214 + src/mSWEEP.cpp:346).
void bb_parameters(uint64_t group_size, double q, double e, double *alpha, double *beta) {
  const double n = (double)group_size;
  const double mean = n * q;
  const double phi = 1.0 / (n - mean + e);
  *beta = phi * (n - mean);
  *alpha = (mean * (*beta)) / (n - mean);
}

// Likelihood.hpp:109-195.
Likelihood build_likelihood(const EcTable &ecs, const Grouping &grouping, double q, double e,
                            double zero_inflation, uint64_t min_hits, bool keep_hit_counts) {
  Likelihood L;
  const uint32_t K = grouping.n_groups();
  const uint64_t N = ecs.n_ecs();
  L.n_groups_all = K;
  L.n_ecs = N;

  // :122-139  c(g, i) = #targets of EC i that belong to group g
  std::vector<uint32_t> counts((size_t)K * N, 0);
#pragma omp parallel for schedule(static)
  for (uint64_t i = 0; i < N; ++i)
    for (uint64_t p = ecs.pat_ptr[i]; p < ecs.pat_ptr[i + 1]; ++p)
      counts[(size_t)grouping.group_of_target[ecs.pat_targets[p]] * N + i] += 1;

  // :141-171  --min-hits mask and compaction
  const bool mask_groups = min_hits > 0;
  L.groups_mask.assign(K, mask_groups ? 0 : 1);
  std::vector<uint32_t> pos(K, 0);
  if (mask_groups) {
    L.group_hits.assign(K, 0);
#pragma omp parallel for schedule(static)
    for (uint32_t g = 0; g < K; ++g) {
      uint64_t h = 0;
      for (uint64_t i = 0; i < N; ++i) h += (uint64_t)(counts[(size_t)g * N + i] > 0) * ecs.count[i];
      L.group_hits[g] = h;
    }
    for (uint32_t g = 0; g < K; ++g) {
      if (L.group_hits[g] >= min_hits) {
        L.groups_mask[g] = 1;
        pos[g] = (uint32_t)L.masked_sizes.size();
        L.masked_sizes.push_back(grouping.sizes[g]);
      }
    }
  } else {
    L.masked_sizes = grouping.sizes;
    std::iota(pos.begin(), pos.end(), 0u);
  }
  const uint32_t Km = (uint32_t)L.masked_sizes.size();
  L.n_groups = Km;

  // :92-107  LUT[g][0] = log(zi); LUT[g][c] = ldbb_scaled(c, n_g, a_g, b_g) + log1p(-zi)
  uint64_t max_size = 0;
  for (uint64_t s : L.masked_sizes) max_size = std::max(max_size, s);
  L.lut_cols = max_size + 1;
  L.lut.assign((size_t)Km * L.lut_cols, std::numeric_limits<double>::quiet_NaN());   // c > n_g is never read
  for (uint32_t g = 0; g < Km; ++g) {
    double a, b;
    bb_parameters(L.masked_sizes[g], q, e, &a, &b);
    L.lut[(size_t)g * L.lut_cols] = std::log(zero_inflation);
    for (uint64_t c = 1; c <= L.masked_sizes[g]; ++c)
      L.lut[(size_t)g * L.lut_cols + c] = ldbb_scaled(c, L.masked_sizes[g], a, b) + std::log1p(-zero_inflation);
  }

  // :176-185  gather
  L.logl.assign((size_t)Km * N, std::log(zero_inflation));
#pragma omp parallel for schedule(static)
  for (uint32_t g = 0; g < K; ++g) {
    if (!L.groups_mask[g]) continue;
    const double *lut_row = &L.lut[(size_t)pos[g] * L.lut_cols];
    for (uint64_t i = 0; i < N; ++i) L.logl[(size_t)pos[g] * N + i] = lut_row[counts[(size_t)g * N + i]];
  }

  // :188-195
  L.log_counts.resize(N);
  for (uint64_t i = 0; i < N; ++i) L.log_counts[i] = std::log((double)ecs.count[i]);

  if (keep_hit_counts) L.hit_counts.swap(counts);
  return L;
}

// =============================================================================================
// Optimiser (rcgpar v1.2.1 restatement — PARITY UNPINNED)
// =============================================================================================

// The authors' digamma: recurrence up to x >= 7, then an asymptotic series in 1/(x - 1/2).
// An in-tree copy of the same function is src/Sample.cpp:87-97.
double digamma_series(double x) {
  double acc = 0.0;
  while (x < 7.0) { acc -= 1.0 / x; x += 1.0; }
  x -= 0.5;
  const double r = 1.0 / x, r2 = r * r, r4 = r2 * r2;
  acc += std::log(x) + (1.0 / 24.0) * r2 - (7.0 / 960.0) * r4 + (31.0 / 8064.0) * r4 * r2 - (127.0 / 30720.0) * r4 * r4;
  return acc;
}

namespace {

struct Stopwatch {
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  double seconds() const { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};

constexpr uint64_t kColBlock = 512;   // columns handled together by one thread (cache blocking)

// gamma(., j) -= logsumexp_k gamma(k, j); lse[j] receives the subtracted value ("oldm").
void normalise_columns(double *gamma, uint32_t K, uint64_t N, double *lse) {
#pragma omp parallel for schedule(static)
  for (uint64_t j0 = 0; j0 < N; j0 += kColBlock) {
    const uint64_t j1 = std::min(N, j0 + kColBlock);
    double m[kColBlock], s[kColBlock];
    for (uint64_t j = j0; j < j1; ++j) { m[j - j0] = -std::numeric_limits<double>::infinity(); s[j - j0] = 0.0; }
    for (uint32_t k = 0; k < K; ++k) {
      const double *g = gamma + (size_t)k * N;
      for (uint64_t j = j0; j < j1; ++j) m[j - j0] = std::max(m[j - j0], g[j]);
    }
    for (uint32_t k = 0; k < K; ++k) {
      const double *g = gamma + (size_t)k * N;
      for (uint64_t j = j0; j < j1; ++j) s[j - j0] += std::exp(g[j] - m[j - j0]);
    }
    for (uint64_t j = j0; j < j1; ++j) lse[j] = m[j - j0] + std::log(s[j - j0]);
    for (uint32_t k = 0; k < K; ++k) {
      double *g = gamma + (size_t)k * N;
      for (uint64_t j = j0; j < j1; ++j) g[j] -= lse[j];
    }
  }
}

// N_k = alpha0_k + sum_j exp(gamma_kj + log c_j)
void update_Nk(const double *gamma, uint32_t K, uint64_t N, const double *log_counts, const double *alpha0, double *N_k) {
#pragma omp parallel for schedule(static)
  for (uint32_t k = 0; k < K; ++k) {
    const double *g = gamma + (size_t)k * N;
    long double acc = 0.0L;
    for (uint64_t j = 0; j < N; ++j) acc += std::exp(g[j] + log_counts[j]);
    N_k[k] = (double)acc + alpha0[k];
  }
}

// ELBO = sum_kj exp(gamma_kj + log c_j) (logl_kj - gamma_kj) + sum_k lgamma(N_k) + const
long double elbo(const double *logl, const double *gamma, uint32_t K, uint64_t N, const double *log_counts,
                 const double *N_k, long double bound_const) {
  long double total = 0.0L;
#pragma omp parallel for schedule(static) reduction(+ : total)
  for (uint32_t k = 0; k < K; ++k) {
    const double *g = gamma + (size_t)k * N;
    const double *l = logl + (size_t)k * N;
    long double acc = 0.0L;
    for (uint64_t j = 0; j < N; ++j) {
      const double w = std::exp(g[j] + log_counts[j]);
      if (w > 0.0) acc += w * (l[j] - g[j]);   // a zero-count EC contributes nothing (bootstrap, -inf log-count)
    }
    total += acc + std::lgamma(N_k[k]);
  }
  return total + bound_const;
}

long double bound_constant(const double *log_counts, uint64_t N, const double *alpha0, uint32_t K) {
  long double n_total = 0.0L, a_total = 0.0L, lg_total = 0.0L;
  for (uint64_t j = 0; j < N; ++j) n_total += std::exp(log_counts[j]);
  for (uint32_t k = 0; k < K; ++k) { a_total += alpha0[k]; lg_total += std::lgamma(alpha0[k]); }
  return std::lgamma((double)a_total) - std::lgamma((double)(a_total + n_total)) - lg_total;
}

// "negative natural gradient": step_kj = logl_kj + digamma(N_k) - 1 - gamma_kj,
// returns sum_kj q_kj (step_kj - <step>_j) step_kj with <step>_j = sum_k q_kj step_kj, q = exp(gamma).
double negnatgrad(const double *gamma, const double *N_k, const double *logl, uint32_t K, uint64_t N, double *step) {
  std::vector<double> dg(K);
  for (uint32_t k = 0; k < K; ++k) dg[k] = digamma_series(N_k[k]) - 1.0;
  long double total = 0.0L;
#pragma omp parallel for schedule(static) reduction(+ : total)
  for (uint64_t j0 = 0; j0 < N; j0 += kColBlock) {
    const uint64_t j1 = std::min(N, j0 + kColBlock);
    double mean[kColBlock];
    for (uint64_t j = j0; j < j1; ++j) mean[j - j0] = 0.0;
    for (uint32_t k = 0; k < K; ++k) {
      const size_t off = (size_t)k * N;
      for (uint64_t j = j0; j < j1; ++j) {
        const double d = logl[off + j] + dg[k] - gamma[off + j];
        step[off + j] = d;
        mean[j - j0] += d * std::exp(gamma[off + j]);
      }
    }
    long double acc = 0.0L;
    for (uint32_t k = 0; k < K; ++k) {
      const size_t off = (size_t)k * N;
      for (uint64_t j = j0; j < j1; ++j)
        acc += std::exp(gamma[off + j]) * (step[off + j] - mean[j - j0]) * step[off + j];
    }
    total += acc;
  }
  return (double)total;
}

} // namespace

// rcgpar::rcg_optl_omp as recalled in SURVEY.md §8(a) (Riemannian conjugate gradient with
// Fletcher-Reeves momentum, restart when the bound decreases, stop on bound change < tol).
ViResult rcg_optl(const double *logl, uint32_t K, uint64_t N, const double *log_counts,
                  const double *alpha0, double tol, uint64_t max_iters) {
  const size_t KN = (size_t)K * N;
  const Stopwatch clock;
  ViResult res;
  res.gamma.assign(KN, std::log(1.0 / (double)K));
  std::vector<double> step(KN, 0.0), oldstep(KN, 0.0), oldm(N, 0.0);
  res.N_k.assign(K, 0.0);
  double *gamma = res.gamma.data();
  double oldnorm = 1.0;
  long double bound = -100000.0L;
  bool didreset = false;
  const long double bconst = bound_constant(log_counts, N, alpha0, K);
  update_Nk(gamma, K, N, log_counts, alpha0, res.N_k.data());

  for (uint64_t it = 0; it < max_iters; ++it) {
    const double newnorm = negnatgrad(gamma, res.N_k.data(), logl, K, N, step.data());
    const double beta_fr = newnorm / oldnorm;
    oldnorm = newnorm;
    if (didreset) {
      std::fill(oldstep.begin(), oldstep.end(), 0.0);
    } else if (beta_fr > 0) {
#pragma omp parallel for schedule(static)
      for (size_t a = 0; a < KN; ++a) { oldstep[a] *= beta_fr; step[a] += oldstep[a]; }
    }
    didreset = false;
#pragma omp parallel for schedule(static)
    for (size_t a = 0; a < KN; ++a) gamma[a] += step[a];
    normalise_columns(gamma, K, N, oldm.data());
    update_Nk(gamma, K, N, log_counts, alpha0, res.N_k.data());
    const long double oldbound = bound;
    bound = elbo(logl, gamma, K, N, log_counts, res.N_k.data(), bconst);

    if (bound < oldbound) {   // the momentum term made things worse: drop it and take the plain step
      didreset = true;
#pragma omp parallel for schedule(static)
      for (uint32_t k = 0; k < K; ++k)
        for (uint64_t j = 0; j < N; ++j) gamma[(size_t)k * N + j] += oldm[j];
      if (beta_fr > 0) {
#pragma omp parallel for schedule(static)
        for (size_t a = 0; a < KN; ++a) gamma[a] -= oldstep[a];
      }
      normalise_columns(gamma, K, N, oldm.data());
      update_Nk(gamma, K, N, log_counts, alpha0, res.N_k.data());
      bound = elbo(logl, gamma, K, N, log_counts, res.N_k.data(), bconst);
    } else {
      oldstep.swap(step);
    }
    res.trace.bound.push_back((double)bound);
    res.trace.gnorm.push_back(newnorm);
    res.trace.reset.push_back(didreset ? 1 : 0);
    res.trace.t_end.push_back(clock.seconds());
    res.iters = it + 1;
    if (bound - oldbound < tol && !didreset) { res.converged = true; break; }
  }
  normalise_columns(gamma, K, N, oldm.data());
  res.bound = (double)bound;
  return res;
}

// rcgpar::em_torch, double precision: coordinate-ascent VB for the same model,
// gamma_kj ∝ exp(logl_kj + digamma(N_k)).  Stops when the bound changes by less than tol
// (first comparison after the second evaluation of the bound).
ViResult em_optl(const double *logl, uint32_t K, uint64_t N, const double *log_counts,
                 const double *alpha0, double tol, uint64_t max_iters) {
  const size_t KN = (size_t)K * N;
  const Stopwatch clock;
  ViResult res;
  res.gamma.assign(KN, std::log(1.0 / (double)K));
  res.N_k.assign(K, 0.0);
  std::vector<double> lse(N, 0.0), dg(K);
  double *gamma = res.gamma.data();
  const long double bconst = bound_constant(log_counts, N, alpha0, K);
  update_Nk(gamma, K, N, log_counts, alpha0, res.N_k.data());
  long double bound = 0.0L;
  for (uint64_t it = 0; it < max_iters; ++it) {
    for (uint32_t k = 0; k < K; ++k) dg[k] = digamma_series(res.N_k[k]);
#pragma omp parallel for schedule(static)
    for (uint32_t k = 0; k < K; ++k)
      for (uint64_t j = 0; j < N; ++j) gamma[(size_t)k * N + j] = logl[(size_t)k * N + j] + dg[k];
    normalise_columns(gamma, K, N, lse.data());
    update_Nk(gamma, K, N, log_counts, alpha0, res.N_k.data());
    const long double oldbound = bound;
    bound = elbo(logl, gamma, K, N, log_counts, res.N_k.data(), bconst);
    res.trace.bound.push_back((double)bound);
    res.trace.gnorm.push_back(0.0);
    res.trace.reset.push_back(0);
    res.trace.t_end.push_back(clock.seconds());
    res.iters = it + 1;
    if (it > 0 && std::fabs((double)(bound - oldbound)) < tol) { res.converged = true; break; }
  }
  res.bound = (double)bound;
  return res;
}

// rcgpar::mixture_components: theta_k = sum_j exp(gamma_kj + log c_j) / sum_j c_j
std::vector<double> mixture_components(const double *gamma, uint32_t K, uint64_t N, const double *log_counts) {
  long double n_total = 0.0L;
  for (uint64_t j = 0; j < N; ++j) n_total += std::exp(log_counts[j]);
  std::vector<double> theta(K, 0.0);
#pragma omp parallel for schedule(static)
  for (uint32_t k = 0; k < K; ++k) {
    long double acc = 0.0L;
    for (uint64_t j = 0; j < N; ++j) acc += std::exp(gamma[(size_t)k * N + j] + log_counts[j]);
    theta[k] = (double)(acc / n_total);
  }
  return theta;
}

// Sample::dirichlet_kld (src/Sample.cpp:99-131): alphas[k] adds exp(gamma_kj) once per read of class j (the read
// count recovered as round(exp(log c_j))), then the KL divergence of the Dirichlet marginal of group k, floored
// at 1e-16.  Sample::get_rates (:133-151): softmax of the log-KLDs with the running maximum started at 0.
RateResult dirichlet_kld(const double *gamma, uint32_t K, uint64_t N, const double *log_counts) {
  std::vector<double> alphas(K, 0.0);
  for (uint32_t k = 0; k < K; ++k)
    for (uint64_t j = 0; j < N; ++j) {
      const size_t reads_in_class = (size_t)std::round(std::exp(log_counts[j]));
      for (size_t r = 0; r < reads_in_class; ++r) alphas[k] += std::exp(gamma[(size_t)k * N + j]);
    }
  double alpha0 = 0.0;
  for (uint32_t k = 0; k < K; ++k) alpha0 += alphas[k];
  RateResult out;
  out.log_kld.resize(K);
  out.rate.resize(K);
  for (uint32_t k = 0; k < K; ++k) {
    const double a = alphas[k];
    const double kld = std::lgamma(alpha0) - std::lgamma(alpha0 - a) - std::lgamma(a) + a * (digamma_series(a) - digamma_series(alpha0));
    out.log_kld[k] = std::log(std::max(kld, 1e-16));
  }
  double top = 0.0;
  for (uint32_t k = 0; k < K; ++k) top = top > out.log_kld[k] ? top : out.log_kld[k];
  double total = 0.0;
  for (uint32_t k = 0; k < K; ++k) total += std::exp(out.log_kld[k] - top);
  const double log_total = std::log(total) + top;
  for (uint32_t k = 0; k < K; ++k) out.rate[k] = std::exp(out.log_kld[k] - log_total);
  return out;
}

std::vector<std::vector<uint32_t>> bin_reads(const double *gamma, uint32_t K, uint64_t N, const std::vector<double> &theta,
                                             const std::vector<uint8_t> &want, const std::vector<uint64_t> &read_ptr,
                                             const std::vector<uint32_t> &read_ids) {
  std::vector<std::vector<uint32_t>> bins(K);
  for (uint32_t k = 0; k < K; ++k) {
    if (!want[k]) continue;
    const double threshold = std::log(theta[k]);
    for (uint64_t j = 0; j < N; ++j)
      if (gamma[(size_t)k * N + j] >= threshold) bins[k].insert(bins[k].end(), read_ids.begin() + read_ptr[j], read_ids.begin() + read_ptr[j + 1]);
    std::sort(bins[k].begin(), bins[k].end());
  }
  return bins;
}

// =============================================================================================
// Bootstrap
// =============================================================================================

struct Bootstrapper::Impl {
  std::mt19937_64 gen;
  std::discrete_distribution<uint32_t> dist;
  uint64_t n_ecs = 0;
};

// src/BootstrapSample.cpp:33-58.  Seed is narrowed to int32_t (include/Sample.hpp:163-169); the
// sentinel 26012023 means "nondeterministic" (:48-53).  Weights are uint32_t EC counts (:38-43).
Bootstrapper::Bootstrapper(const std::vector<uint64_t> &ec_counts, int32_t seed, uint64_t count) : impl(new Impl) {
  if (seed == 26012023) {
    std::random_device rd;
    impl->gen = std::mt19937_64(rd());
  } else {
    impl->gen = std::mt19937_64(seed);
  }
  std::vector<uint32_t> weights(ec_counts.size());
  uint64_t total = 0;
  for (size_t i = 0; i < ec_counts.size(); ++i) { weights[i] = (uint32_t)ec_counts[i]; total += ec_counts[i]; }
  impl->dist = std::discrete_distribution<uint32_t>(weights.begin(), weights.end());
  impl->n_ecs = ec_counts.size();
  bootstrap_count = count == 0 ? total : count;   // :56
}
Bootstrapper::~Bootstrapper() { delete impl; }

// src/BootstrapSample.cpp:60-66
std::vector<uint32_t> Bootstrapper::resample_raw() {
  std::vector<uint32_t> tmp(impl->n_ecs, 0);
  for (uint64_t i = 0; i < bootstrap_count; ++i) tmp[impl->dist(impl->gen)] += 1;
  return tmp;
}

// src/BootstrapSample.cpp:67-72
std::vector<double> Bootstrapper::resample_counts() {
  const std::vector<uint32_t> tmp = resample_raw();
  std::vector<double> out(tmp.size());
  for (size_t i = 0; i < tmp.size(); ++i) out[i] = std::log((double)tmp[i]);
  return out;
}

// =============================================================================================
// Output
// =============================================================================================

// src/PlainSample.cpp:32-71 and src/BootstrapSample.cpp:75-130 (default ostream formatting).
void write_abundances(std::ostream &of, const std::string &version, uint64_t n_reads, uint64_t n_aligned,
                      const std::vector<std::string> &estimated_names, const std::vector<std::string> &zero_names,
                      const std::vector<std::vector<double>> &results, uint64_t bootstrap_iters) {
  if (!of.good()) throw std::runtime_error("Can't write to abundances file.");
  of << "#mSWEEP_version:" << '\t' << version << '\n';
  of << "#num_reads:" << '\t' << n_reads << '\n';
  of << "#num_aligned:" << '\t' << n_aligned << '\n';
  if (bootstrap_iters > 0) {
    of << "#bootstrap_iters:" << '\t' << bootstrap_iters << '\n';
    of << "#c_id" << '\t' << "mean_theta" << '\t' << "bootstrap_mean_thetas" << '\n';
  } else {
    of << "#c_id" << '\t' << "mean_theta" << '\n';
  }
  const size_t n_est = estimated_names.size();
  for (size_t i = 0; i < n_est + zero_names.size(); ++i) {
    const bool est = i < n_est;
    of << (est ? estimated_names[i] : zero_names[i - n_est]) << '\t';
    of << (est ? results[0][i] : 0.0);
    for (uint64_t b = 0; b < bootstrap_iters; ++b) of << '\t' << (est ? results[b + 1][i] : 0.0);
    of << '\n';
  }
  of.flush();
}

} // namespace oracle
