// oracle/oracle_main.cpp — command-line front end of the CPU oracle with the reference's flags for the
// abundance-estimation path (src/mSWEEP.cpp:68-160).  TEST INFRASTRUCTURE ONLY: lets the tests compare the
// product's mSWEEP_b200 binary against the restated reference end to end, file against file.
#include "oracle.hpp"

#include <chrono>
#include <cmath>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>

#ifdef _OPENMP
#include <omp.h>
#endif

using namespace oracle;

int main(int argc, char **argv) {
  std::map<std::string, std::string> kv;
  for (int i = 1; i < argc; ++i) {
    std::string k = argv[i];
    k = k.substr(k[1] == '-' ? 2 : 1);
    if (k == "verbose") continue;
    if (k == "write-probs" || k == "run-rate" || k == "bin-reads" || k == "print-timings") { kv[k] = "1"; continue; }
    if (i + 1 >= argc) { std::cerr << "missing value for " << k << "\n"; return 1; }
    kv[k] = argv[++i];
  }
  auto get = [&](const std::string &k, const std::string &d) { return kv.count(k) ? kv[k] : d; };
  // stage timers for the CPU baseline: parse, collapse, likelihood, optimiser, bootstrap, write
  auto t_last = std::chrono::steady_clock::now();
  auto lap = [&]() { const auto t = std::chrono::steady_clock::now(); const double s = std::chrono::duration<double>(t - t_last).count(); t_last = t; return s; };
  double t_parse = 0, t_collapse = 0, t_lik = 0, t_vi = 0, t_boot = 0;
  uint64_t vi_iters = 0;
  try {
#ifdef _OPENMP
    omp_set_num_threads(std::stoi(get("t", "1")));
#endif
    std::ifstream gi(get("i", ""));
    Grouping grouping = read_grouping(gi);
    std::vector<std::string> paths;
    if (kv.count("themisto")) { std::stringstream ss(kv["themisto"]); std::string p; while (std::getline(ss, p, ',')) paths.push_back(p); }
    else paths = {get("themisto-1", ""), get("themisto-2", "")};
    std::vector<std::ifstream> files;
    for (auto &p : paths) { files.emplace_back(p); if (!files.back().good()) throw std::runtime_error("cannot open " + p); }
    std::vector<std::istream*> strands;
    for (auto &f : files) strands.push_back(&f);
    ReadTable reads = read_themisto(strands, grouping.group_of_target.size(), get("themisto-mode", "intersection"));
    t_parse = lap();
    EcTable ec = collapse(reads);
    t_collapse = lap();
    const uint64_t min_hits = std::stoull(get("min-hits", "0"));
    Likelihood lik = build_likelihood(ec, grouping, std::stod(get("q", "0.65")), std::stod(get("e", "0.01")),
                                      std::stod(get("zero-inflation", "0.01")), min_hits, false);
    t_lik = lap();
    const uint32_t K = lik.n_groups;
    std::vector<double> prior(K, 1.0);
    if (kv.count("alphas")) { prior.clear(); std::stringstream ss(kv["alphas"]); std::string p; while (std::getline(ss, p, ',')) prior.push_back(std::stod(p)); }
    const double tol = std::stod(get("tol", "0.000001"));
    const uint64_t max_iters = std::stoull(get("max-iters", "5000"));
    const std::string algo = get("algorithm", "rcgcpu");
    const bool rcg = algo.rfind("rcg", 0) == 0;
    std::vector<double> first_gamma;
    auto estimate = [&](const std::vector<double> &lc) {
      ViResult r = rcg ? rcg_optl(lik.logl.data(), K, ec.n_ecs(), lc.data(), prior.data(), tol, max_iters)
                       : em_optl(lik.logl.data(), K, ec.n_ecs(), lc.data(), prior.data(), tol, max_iters);
      if (first_gamma.empty()) { first_gamma = r.gamma; vi_iters = r.iters; }
      return mixture_components(r.gamma.data(), K, ec.n_ecs(), lc.data());
    };
    std::vector<std::vector<double>> results;
    results.push_back(estimate(lik.log_counts));
    t_vi = lap();
    const uint64_t iters = std::stoull(get("iters", "0"));
    if (iters) {
      Bootstrapper bs(ec.count, (int32_t)std::stoull(get("seed", "26012023")), std::stoull(get("bootstrap-count", "0")));
      for (uint64_t r = 0; r < iters; ++r) results.push_back(estimate(bs.resample_counts()));
    }
    t_boot = lap();
    std::vector<std::string> est, zero;
    for (size_t g = 0; g < grouping.names.size(); ++g) (lik.groups_mask[g] ? est : zero).push_back(grouping.names[g]);
    uint64_t n_aligned = 0;
    for (auto c : ec.count) n_aligned += c;
    if (kv.count("write-probs")) {   // src/Sample.cpp:63-85, 154-186
      std::ofstream pf(get("o", "oracle") + "_probs.tsv");
      pf << "ec_id" << '\t';
      const size_t n_rows = est.size() + zero.size();
      for (size_t i = 0; i < n_rows; ++i) pf << (i < est.size() ? est[i] : zero[i - est.size()]) << (i + 1 < n_rows ? '\t' : '\n');
      for (uint64_t j = 0; j < ec.n_ecs(); ++j) {
        pf << j << '\t';
        for (size_t i = 0; i < n_rows; ++i)
          pf << (i < est.size() ? std::exp(first_gamma[i * ec.n_ecs() + j]) : 0.0) << (i + 1 < n_rows ? '\t' : '\n');
      }
      pf << std::endl;
    }
    if (kv.count("bin-reads")) {   // src/mSWEEP.cpp:437-469, src/OutfileDesignator.cpp:80-94
      std::vector<uint8_t> want(K, kv.count("target-groups") ? 0 : 1);
      if (kv.count("target-groups")) {
        std::stringstream ss(kv["target-groups"]); std::string name;
        while (std::getline(ss, name, ',')) for (uint32_t k = 0; k < K; ++k) if (est[k] == name) want[k] = 1;
      }
      if (kv.count("min-abundance")) for (uint32_t k = 0; k < K; ++k) if (results[0][k] < std::stod(kv["min-abundance"])) want[k] = 0;
      const auto bins = bin_reads(first_gamma.data(), K, ec.n_ecs(), results[0], want, ec.read_ptr, ec.read_ids);
      std::string dir = ".";
      const std::string o = get("o", "oracle");
      if (o.find('/') != std::string::npos) dir = o.substr(0, o.rfind('/'));
      for (uint32_t k = 0; k < K; ++k) {
        if (!want[k]) continue;
        std::ofstream bf(dir + '/' + est[k] + ".bin");
        for (uint32_t r : bins[k]) bf << (uint64_t)r + 1 << '\n';   // 1-based, as mGEMS extract counts reads (assumption)
      }
    }
    std::ofstream of(get("o", "oracle") + "_abundances.txt");
    if (kv.count("run-rate")) {    // src/mSWEEP.cpp:524-548
      const RateResult rr = dirichlet_kld(first_gamma.data(), K, ec.n_ecs(), lik.log_counts.data());
      of << "#mSWEEP_version:" << '\t' << get("version-string", "oracle") << '\n';
      of << "#num_reads:" << '\t' << reads.n_reads << '\n';
      of << "#num_aligned:" << '\t' << n_aligned << '\n';
      of << "#c_id" << '\t' << "mean_theta" << '\t' << "RATE" << '\t' << "KLD" << '\n';
      for (size_t i = 0; i < est.size(); ++i) of << est[i] << '\t' << results[0][i] << '\t' << rr.rate[i] << '\t' << std::exp(rr.log_kld[i]) << '\n';
      for (size_t i = 0; i < zero.size(); ++i) of << zero[i] << '\t' << 0.0 << '\t' << 0.0 << '\t' << 0.0 << '\n';
    } else
    write_abundances(of, get("version-string", "oracle"), reads.n_reads, n_aligned, est, zero, results, iters);
    if (kv.count("print-timings"))
      std::cerr << "{\"parse_s\": " << t_parse << ", \"collapse_s\": " << t_collapse << ", \"likelihood_s\": " << t_lik
                << ", \"optimiser_s\": " << t_vi << ", \"optimiser_iters\": " << vi_iters
                << ", \"optimiser_s_per_iter\": " << (vi_iters ? t_vi / (double)vi_iters : 0.0) << ", \"bootstrap_s\": " << t_boot
                << ", \"write_s\": " << lap() << ", \"n_ecs\": " << ec.n_ecs() << ", \"threads\": " << get("t", "1") << "}" << std::endl;
  } catch (const std::exception &e) {
    std::cerr << "oracle failed:\n  " << e.what() << "\nexiting\n";
    return 1;
  }
  return 0;
}
