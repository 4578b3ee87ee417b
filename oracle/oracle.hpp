// oracle/oracle.hpp — CPU restatement of mSWEEP's abundance-estimation hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product (msweep_b200/, include/) may include,
// link or call this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` leg use it, and only as the checker / CPU baseline.
//
// Parity status:
//   * EC construction, hit counts, --min-hits mask, LUT, bootstrap resampling: PINNED by the
//     in-tree reference sources cited at each function (integer semantics, libstdc++ <random>).
//   * Optimiser (rcg_optl_omp / em / mixture_components): PARITY UNPINNED.  The arithmetic lives
//     in github.com/tmaklin/rcgpar @ v1.2.1 (reference CMakeLists.txt:280-282), which is not in
//     /root/reference and cannot be fetched; the reference ships no tests or golden vectors.
//     The functions here restate the published algorithm (Mäklin et al. 2021, Wellcome Open Res
//     5:14) anchored on the reference's call sites src/mSWEEP.cpp:176-205, 419-423, 507-516.
//
// All file:line citations are relative to /root/reference/.
#pragma once
#include <cstdint>
#include <cstddef>
#include <string>
#include <vector>
#include <istream>

namespace oracle {

// ---------------------------------------------------------------------------------------------
// Pseudoalignment (include/mSWEEP_alignment.hpp:54-135)
// ---------------------------------------------------------------------------------------------
struct ReadTable {
  uint64_t n_reads = 0;     // reference `n_queries`: line count of the LAST strand (:121)
  uint64_t n_targets = 0;
  // rows[i] = ascending, unique target ids with bit (i, t) set after strand merging, i < n_reads
  std::vector<std::vector<uint32_t>> rows;
};

// Parses one Themisto plaintext stream into a set of flat bit indices read_id*T + target
// (ReadPlaintextLine, :54-66).  Returns the number of lines.
uint64_t parse_plaintext_strand(std::istream &in, uint64_t n_targets, std::vector<uint64_t> *bits);

// read() (:97-135): strands merged by "intersection" (bit_and) or "union" (bit_or).
ReadTable read_themisto(const std::vector<std::istream*> &strands, uint64_t n_targets,
                        const std::string &merge_mode);

ReadTable from_csr(uint64_t n_reads, uint64_t n_targets, const uint64_t *row_ptr, const uint32_t *targets);

// ---------------------------------------------------------------------------------------------
// Equivalence classes (Alignment::collapse, include/mSWEEP_alignment.hpp:137-215)
// ---------------------------------------------------------------------------------------------
uint64_t pattern_hash(const uint32_t *targets, size_t n);   // :150-155

struct EcTable {
  uint64_t n_reads = 0, n_targets = 0;
  std::vector<uint64_t> hash;       // ascending (std::map order, :200)
  std::vector<uint64_t> count;      // ec_counts (:203)
  std::vector<uint32_t> rep_read;   // ec_read_ids[i][0], the smallest read id of the class (:204-206)
  std::vector<uint64_t> read_ptr;   // CSR over read_ids
  std::vector<uint32_t> read_ids;   // ascending inside each class
  std::vector<uint64_t> pat_ptr;    // CSR over representative patterns (the collapsed `bits`)
  std::vector<uint32_t> pat_targets;
  uint64_t n_ecs() const { return hash.size(); }
};
EcTable collapse(const ReadTable &reads);

// ---------------------------------------------------------------------------------------------
// Grouping (include/Grouping.hpp:62-83, include/Reference.hpp:67-94, src/Reference.cpp:31-56)
// ---------------------------------------------------------------------------------------------
struct Grouping {
  std::vector<std::string> names;         // order of first appearance
  std::vector<uint64_t> sizes;
  std::vector<uint32_t> group_of_target;  // one per indicator line
  uint32_t n_groups() const { return (uint32_t)names.size(); }
};
Grouping read_grouping(std::istream &in, char delimiter = '\t', size_t column = 0);

// ---------------------------------------------------------------------------------------------
// Likelihood (include/Likelihood.hpp:47-60, 92-107, 109-207)
// ---------------------------------------------------------------------------------------------
double lbeta(double x, double y);
double ldbb_scaled(uint64_t k, uint64_t n, double alpha, double beta);
void bb_parameters(uint64_t group_size, double q, double e, double *alpha, double *beta);   // :198-207

struct Likelihood {
  uint32_t n_groups_all = 0;          // K before masking
  uint32_t n_groups = 0;              // K' rows of the matrix
  uint64_t n_ecs = 0;
  std::vector<uint8_t> groups_mask;   // groups_considered() (:157, :142)
  std::vector<uint64_t> group_hits;   // hits[g] (only when min_hits > 0)
  std::vector<uint64_t> masked_sizes; // sizes of kept groups
  std::vector<uint32_t> hit_counts;   // c(g, i): n_groups_all x n_ecs, group-major (optional)
  uint64_t lut_cols = 0;              // max kept group size + 1
  std::vector<double> lut;            // K' x lut_cols (precalc_lls, :92-107)
  std::vector<double> logl;           // K' x n_ecs group-major (:176-185)
  std::vector<double> log_counts;     // log(ec_count) (:188-195)
};
Likelihood build_likelihood(const EcTable &ecs, const Grouping &grouping, double q, double e,
                            double zero_inflation, uint64_t min_hits, bool keep_hit_counts);

// ---------------------------------------------------------------------------------------------
// Optimiser — rcgpar v1.2.1 restatement (PARITY UNPINNED, see header)
// ---------------------------------------------------------------------------------------------
double digamma_series(double x);   // same series as src/Sample.cpp:87-97

struct ViTrace { std::vector<double> bound, gnorm, t_end /* seconds since entry, per iteration */; std::vector<uint8_t> reset; };
struct ViResult {
  std::vector<double> gamma;   // K x N group-major log-posteriors (what rcg_optl_* returns)
  std::vector<double> N_k;     // alpha0 + expected counts at exit
  double bound = 0;
  uint64_t iters = 0;          // iterations executed
  bool converged = false;
  ViTrace trace;
};

// rcgpar::rcg_optl_omp (called at src/mSWEEP.cpp:198)
ViResult rcg_optl(const double *logl, uint32_t K, uint64_t N, const double *log_counts,
                  const double *alpha0, double tol, uint64_t max_iters);
// rcgpar::em_torch in double precision (called at src/mSWEEP.cpp:202)
ViResult em_optl(const double *logl, uint32_t K, uint64_t N, const double *log_counts,
                 const double *alpha0, double tol, uint64_t max_iters);
// rcgpar::mixture_components (called at src/mSWEEP.cpp:420)
std::vector<double> mixture_components(const double *gamma, uint32_t K, uint64_t N, const double *log_counts);

// --run-rate: Sample::dirichlet_kld and Sample::get_rates (src/Sample.cpp:99-151).  PINNED by the in-tree code.
struct RateResult { std::vector<double> log_kld, rate; };
RateResult dirichlet_kld(const double *gamma, uint32_t K, uint64_t N, const double *log_counts);

// --bin-reads: the rule of mGEMS::BinFromMatrix as called at src/mSWEEP.cpp:437-469 (mGEMS v1.3.3 is OFF-TREE,
// CMakeLists.txt:318-320: PARITY UNPINNED, restated from the published description).  bins[k] = ascending ids of
// the reads whose class has gamma(k, j) >= log(theta[k]), for the groups with want[k] != 0.
std::vector<std::vector<uint32_t>> bin_reads(const double *gamma, uint32_t K, uint64_t N, const std::vector<double> &theta,
                                             const std::vector<uint8_t> &want, const std::vector<uint64_t> &read_ptr,
                                             const std::vector<uint32_t> &read_ids);

// ---------------------------------------------------------------------------------------------
// Bootstrap (src/BootstrapSample.cpp:33-73, include/Sample.hpp:163-174)
// ---------------------------------------------------------------------------------------------
struct Bootstrapper {
  Bootstrapper(const std::vector<uint64_t> &ec_counts, int32_t seed, uint64_t bootstrap_count);
  ~Bootstrapper();
  std::vector<uint32_t> resample_raw();    // tmp_counts of resample_counts() (:61-66)
  std::vector<double> resample_counts();   // log(tmp_counts) with -inf for zero (:67-72)
  uint64_t bootstrap_count;
  struct Impl; Impl *impl;
};

// Output (src/PlainSample.cpp:32-71, src/BootstrapSample.cpp:75-130)
void write_abundances(std::ostream &of, const std::string &version, uint64_t n_reads, uint64_t n_aligned,
                      const std::vector<std::string> &estimated_names, const std::vector<std::string> &zero_names,
                      const std::vector<std::vector<double>> &results /* [0]=plain, [1..]=bootstrap */,
                      uint64_t bootstrap_iters);

} // namespace oracle
