// main.cpp — mSWEEP_b200: the reference's command line for the abundance-estimation path
// (src/mSWEEP.cpp:68-160, 207-567), driving the B200 backend through the C ABI.
//
// Kept from the reference: the flags of the hot path with their defaults and meaning, the order of the
// stages, the messages and exit codes of the failure sites, and the <prefix>_abundances.txt format
// (src/PlainSample.cpp:32-71, src/BootstrapSample.cpp:75-130), --run-rate (src/Sample.cpp:99-151) and the
// --bin-reads hand-off (src/mSWEEP.cpp:437-469).  Not here (out of scope): likelihood dumps,
// output compression, the compact alignment format (gzip input is read).
// New: --algorithm takes the B200 backends (rcgb200 | emb200; the reference's rcggpu / emgpu are accepted
// as aliases and rcgcpu is refused: there is no CPU path in this binary), and --gpus N.
#include "input.hpp"
#include "msweep_b200.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <thread>

#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef MSWEEP_BUILD_VERSION
#define MSWEEP_BUILD_VERSION "b200-0.1.0"
#endif

namespace {

struct Args {
  std::map<std::string, std::string> kv;
  std::set<std::string> flags;
  bool has(const std::string &k) const { return kv.count(k) || flags.count(k); }
  std::string str(const std::string &k, const std::string &dflt) const { auto it = kv.find(k); return it == kv.end() ? dflt : it->second; }
  template <typename T> T num(const std::string &k, T dflt) const {
    auto it = kv.find(k);
    if (it == kv.end()) return dflt;
    std::istringstream ss(it->second);
    T v;
    if (!(ss >> v)) throw std::runtime_error("Could not parse the value of --" + k + ": " + it->second);
    return v;
  }
};

const std::set<std::string> kBool = {"verbose", "version", "cite", "help", "no-fit-model", "print-timings", "write-probs", "print-probs",
                                     "bin-reads", "run-rate"};
const std::set<std::string> kValued = {"themisto-1", "themisto-2", "themisto", "i", "o", "themisto-mode", "t", "max-iters", "tol",
                                       "algorithm", "emprecision", "iters", "seed", "bootstrap-count", "q", "e", "alphas",
                                       "zero-inflation", "min-hits", "gpus", "rng", "dump-alignment", "storage",
                                       "target-groups", "min-abundance"};
const std::set<std::string> kUnsupported = {"write-likelihood", "write-likelihood-bitseq", "compress", "compression-level",
                                            "read-likelihood"};

Args parse(int argc, char **argv) {
  Args a;
  for (int i = 1; i < argc; ++i) {
    std::string tok = argv[i];
    if (tok.size() < 2 || tok[0] != '-') throw std::runtime_error("Unexpected argument: " + tok);
    std::string key = tok.substr(tok[1] == '-' ? 2 : 1), val;
    const size_t eq = key.find('=');
    bool has_val = false;
    if (eq != std::string::npos) { val = key.substr(eq + 1); key = key.substr(0, eq); has_val = true; }
    if (kUnsupported.count(key)) throw std::runtime_error("--" + key + " is outside the scope of the B200 abundance-estimation backend");
    if (kBool.count(key)) { a.flags.insert(key); continue; }
    if (!kValued.count(key)) throw std::runtime_error("Unknown argument: " + tok);
    if (!has_val) {
      if (i + 1 >= argc) throw std::runtime_error("Argument " + tok + " needs a value");
      val = argv[++i];
    }
    a.kv[key] = val;
  }
  return a;
}

std::vector<std::string> split(const std::string &s, char d) {
  std::vector<std::string> out;
  std::stringstream ss(s);
  std::string part;
  while (std::getline(ss, part, d)) out.push_back(part);
  return out;
}

const char *kHelp =
    "Usage: mSWEEP_b200 --themisto-1 <forwardPseudoalignments> --themisto-2 <reversePseudoalignments> -i <groupIndicatorsFile>\n\n"
    "  --themisto-1, --themisto-2   Themisto pseudoalignments of the two strands (plaintext)\n"
    "  --themisto a[,b]             single file or comma separated list\n"
    "  -i                           group indicators, one line per reference sequence (required)\n"
    "  -o                           output prefix (default: print to cout)\n"
    "  --themisto-mode              intersection | union (default: intersection)\n"
    "  -t                           host threads for parsing (default: 1)\n"
    "  --gpus                       number of B200s to use (default: 1)\n"
    "  --algorithm                  rcgb200 | emb200 (aliases: rcggpu, emgpu; default: rcgb200)\n"
    "  --emprecision                float | double, for emb200 (default: double)\n"
    "  --storage                    auto | dense | sparse likelihood on the device (default: auto = sparse, lossless, O(hits) per class)\n"
    "  --max-iters                  optimiser iteration cap (default: 5000)\n"
    "  --tol                        stop when the bound changes by less than this (default: 0.000001)\n"
    "  --iters                      bootstrap replicates (default: 0)\n"
    "  --seed                       bootstrap seed (default: random)\n"
    "  --bootstrap-count            pseudoalignments to resample per replicate (default: number of aligned reads)\n"
    "  --rng                        exact | philox bootstrap generator (default: exact = std::mt19937_64 stream)\n"
    "  -q, -e                       beta-binomial mean and dispersion (defaults: 0.65, 0.01)\n"
    "  --alphas                     comma separated prior counts (default: all 1.0)\n"
    "  --zero-inflation             likelihood of zero hits against a group (default: 0.01)\n"
    "  --min-hits                   only consider groups with at least this many aligned reads (default: 0)\n"
    "  --write-probs, --print-probs write / print the read-class to group probabilities (<prefix>_probs.tsv)\n"
    "  --bin-reads                  assign reads to groups (mGEMS rule) and write <group>.bin files to the directory of -o\n"
    "  --target-groups a[,b]        only bin these groups (default: all estimated groups)\n"
    "  --min-abundance              only bin groups with at least this relative abundance\n"
    "  --run-rate                   add the RATE and KLD columns to the abundances\n"
    "  --no-fit-model, --verbose, --version, --cite, --help, --print-timings\n";

void cite() {
  std::cerr << "Please cite us as:\n"
            << "\tMäklin T, Kallonen T, David S et al. High-resolution sweep\n"
            << "\tmetagenomics using fast probabilistic inference [version 2;\n"
            << "\tpeer review: 2 approved]. Wellcome Open Res 2021, 5:14\n"
            << "\t(https://doi.org/10.12688/wellcomeopenres.15639.2)" << std::endl;
}

// src/PlainSample.cpp:32-71 and src/BootstrapSample.cpp:75-130: default ostream formatting throughout.
void write_abundances(std::ostream &of, uint64_t n_reads, uint64_t n_aligned, const std::vector<std::string> &estimated,
                      const std::vector<std::string> &zero, const std::vector<std::vector<double>> &results, uint64_t iters) {
  if (!of.good()) throw std::runtime_error(iters ? "Could not write to abundances file." : "Can't write to abundances file.");
  of << "#mSWEEP_version:" << '\t' << MSWEEP_BUILD_VERSION << '\n';
  of << "#num_reads:" << '\t' << n_reads << '\n';
  of << "#num_aligned:" << '\t' << n_aligned << '\n';
  if (iters) {
    of << "#bootstrap_iters:" << '\t' << iters << '\n';
    of << "#c_id" << '\t' << "mean_theta" << '\t' << "bootstrap_mean_thetas" << '\n';
  } else {
    of << "#c_id" << '\t' << "mean_theta" << '\n';
  }
  for (size_t i = 0; i < estimated.size() + zero.size(); ++i) {
    const bool est = i < estimated.size();
    of << (est ? estimated[i] : zero[i - estimated.size()]) << '\t' << (est ? results[0][i] : 0.0);
    for (uint64_t b = 0; b < iters; ++b) of << '\t' << (est ? results[b + 1][i] : 0.0);
    of << '\n';
  }
  of.flush();
}

struct Timer {
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  double lap() { auto t1 = std::chrono::steady_clock::now(); double s = std::chrono::duration<double>(t1 - t0).count(); t0 = t1; return s; }
};

} // namespace

int main(int argc, char *argv[]) {
  bool verbose = false;
  for (int i = 1; i < argc; ++i) verbose = verbose || std::strcmp(argv[i], "--verbose") == 0;
  auto log = [&](const std::string &m) { if (verbose) std::cerr << m << '\n'; };
  log(std::string("mSWEEP-") + MSWEEP_BUILD_VERSION + " abundance estimation");

  Args args;
  try {
    log("Parsing arguments");
    args = parse(argc, argv);
    if (args.has("help")) std::cerr << "\n" << kHelp << "\n\n";
    if (args.has("version")) std::cerr << "mSWEEP-" << MSWEEP_BUILD_VERSION << std::endl;
    if (args.has("cite")) cite();
    if (args.has("help") || args.has("version") || args.has("cite")) return 0;
    if (!args.has("i")) throw std::runtime_error("Required argument -i was not given");
    if (args.has("themisto-1") != args.has("themisto-2")) throw std::runtime_error("--themisto-1 and --themisto-2 must be given together");
    if (!args.has("themisto") && !args.has("themisto-1")) throw std::runtime_error("No pseudoalignment files were given (--themisto-1/--themisto-2 or --themisto)");
    const std::string o = args.str("o", "");
    if (o.find('/') != std::string::npos) {
      std::string dir = o.substr(0, o.rfind('/'));
      std::ofstream probe(dir + "/.mSWEEP_b200_probe");
      if (!probe.good()) throw std::runtime_error("Directory " + dir + " does not seem to exist.");
      probe.close();
      std::remove((dir + "/.mSWEEP_b200_probe").c_str());
    }
  } catch (const std::exception &e) {
    std::cerr << "Error in parsing arguments:\n  " << e.what() << "\nexiting\n";
    return 1;
  }

  const int n_threads = (int)std::max<size_t>(1, args.num<size_t>("t", 1));
#ifdef _OPENMP
  omp_set_num_threads(n_threads);
#endif
  const int n_gpus = std::max(1, args.num<int>("gpus", 1));
  const bool timings = args.has("print-timings");
  Timer timer;
  // CUDA context creation takes seconds on a cold process, and the first launch of every kernel loads its code: both
  // can happen on side threads while the inputs are parsed.  Stage 1: the context.  Stage 2 (opt-in, see below), once the
  // grouping and the options are known: a miniature estimate (64 reads) with the run's own group count, storage and
  // algorithm, so that the kernels the real run will launch are resident when the reads arrive.  Best effort: errors are ignored.
  struct WarmPlan { std::mutex mu; std::condition_variable cv; bool ready = false, go = false;
                    std::function<void(int)> run; } warm_plan;
  std::vector<std::thread> warmups;
  if (!args.has("dump-alignment"))
    for (int g = 0; g < n_gpus; ++g) warmups.emplace_back([g, &warm_plan] {
      mswb_device_warmup(g);
      std::unique_lock<std::mutex> lock(warm_plan.mu);
      warm_plan.cv.wait(lock, [&] { return warm_plan.ready; });
      if (!warm_plan.go) return;
      lock.unlock();
      try { warm_plan.run(g); } catch (...) {}
    });
  struct JoinAll { std::vector<std::thread> &t; WarmPlan &p;
                   ~JoinAll() { { std::lock_guard<std::mutex> l(p.mu); p.ready = true; } p.cv.notify_all();
                                for (auto &x : t) if (x.joinable()) x.join(); } } join_warmups{warmups, warm_plan};

  // ---- group indicators (src/mSWEEP.cpp:258-273) ---------------------------------------------------
  b200::Grouping grouping;
  log("Reading the input files");
  try {
    log("  reading group indicators");
    grouping = b200::read_grouping(args.str("i", ""));
    if (grouping.n_groupings > 1)
      throw std::runtime_error("multiple groupings per indicator file are outside the scope of this backend (use a single column)");
    log("  read " + std::to_string(grouping.group_of_target.size()) + " group indicators");
  } catch (std::exception &e) {
    std::cerr << "Reading group indicators failed:\n  " << e.what() << "\nexiting\n";
    return 1;
  }
  const double t_grouping = timer.lap();

  // ---- algorithm selection (src/mSWEEP.cpp:127, 192-203; unknown strings are refused, not run as EM) ---
  b200::ViOptions vi;
  int storage = MSWB_STORE_F64;
  bool sparse_if_pruned = false;   // auto mode, more than 4096 groups: --min-hits may leave few enough (config 4: 10,000 -> ~100)
  try {
    const std::string algo = args.str("algorithm", "rcgb200");
    if (algo == "rcgb200" || algo == "rcggpu") vi.algo = MSWB_ALGO_RCG;
    else if (algo == "emb200" || algo == "emgpu") vi.algo = MSWB_ALGO_EM;
    else if (algo == "rcgcpu") throw std::runtime_error("--algorithm rcgcpu: this binary is the B200 backend and has no CPU path");
    else throw std::runtime_error("Unknown --algorithm `" + algo + "` (one of rcgb200, emb200)");
    const std::string prec = args.str("emprecision", "double");
    if (prec == "float") { if (vi.algo == MSWB_ALGO_EM) storage = MSWB_STORE_F32; }
    else if (prec != "double") throw std::runtime_error("Unknown --emprecision `" + prec + "` (one of float, double)");
    // auto: the lossless sparse form (O(hits) per class instead of O(groups): LL_WOR21 gives every group a class does
    // not hit the same value) whenever the likelihood is fp64; --emprecision float is a dense form by definition.
    const std::string store = args.str("storage", "auto");
    // (the sparse sweeps keep their K-vectors in shared memory: up to 4096 groups; wider groupings stay dense in auto mode)
    if (store == "sparse" || (store == "auto" && storage == MSWB_STORE_F64 && grouping.sizes.size() <= 4096)) storage = MSWB_STORE_SPARSE;
    else if (store == "auto" && storage == MSWB_STORE_F64) sparse_if_pruned = true;
    else if (store != "dense" && store != "auto") throw std::runtime_error("Unknown --storage `" + store + "` (one of auto, dense, sparse)");
    if (store == "sparse" && args.str("emprecision", "double") == "float" && vi.algo == MSWB_ALGO_EM)
      throw std::runtime_error("--storage sparse is an fp64 form; drop --emprecision float");
    vi.tol = args.num<double>("tol", 1e-6);
    vi.max_iters = args.num<uint64_t>("max-iters", 5000);
  } catch (std::exception &e) {
    std::cerr << "Error in parsing arguments:\n  " << e.what() << "\nexiting\n";
    return 1;
  }

  const uint64_t iters = args.num<uint64_t>("iters", 0);
  if (iters > 65535) { std::cerr << "Error in parsing arguments:\n  --iters is limited to 65535 replicates\nexiting\n"; return 1; }   // uint16_t loop, src/mSWEEP.cpp:498
  const bool bootstrap = iters > 0;
  const uint64_t min_hits = args.num<uint64_t>("min-hits", 0);
  const double q = args.num<double>("q", 0.65), e_disp = args.num<double>("e", 0.01), zi = args.num<double>("zero-inflation", 0.01);
  const int32_t seed = (int32_t)args.num<size_t>("seed", 26012023);     // narrowed as in include/Sample.hpp:163-169
  const uint64_t bootstrap_count = args.num<uint64_t>("bootstrap-count", 0);
  const int rng_mode = args.str("rng", "exact") == "philox" ? MSWB_RNG_PHILOX : MSWB_RNG_LIBSTDCXX_EXACT;

  // stage 2 of the warm-up (see above)
  {
    std::lock_guard<std::mutex> l(warm_plan.mu);
    warm_plan.run = [&](int g) {
      const size_t T = grouping.group_of_target.size();
      if (T == 0) return;
      b200::ReadTable tiny;
      tiny.n_reads = 64; tiny.n_targets = T;
      tiny.row_ptr.assign(65, 0);
      tiny.targets.resize(64);
      for (uint64_t r = 0; r < 64; ++r) { tiny.targets[r] = (uint32_t)((r * 7919u) % T); tiny.row_ptr[r + 1] = r + 1; }
      b200::Context ctx(g);
      b200::Alignment aln(ctx, tiny);
      b200::Likelihood ll(ctx, aln, grouping.group_of_target, grouping.sizes, q, e_disp, 0, zi, storage);
      if (args.has("no-fit-model")) return;
      b200::ViOptions o = vi;
      o.max_iters = 2;
      const std::vector<double> prior(ll.get_rows(), 1.0);
      b200::rcg_optl(ctx, ll, nullptr, prior, o);
      if (bootstrap) ll.bootstrap(prior, o, 1, 16, 1, 0, 1, rng_mode);
    };
    // Opt-in (MSWB_WARM=1).  Measured on config 1 (1e6 reads, 0.25 s of parsing): the miniature estimate takes ~1.2 s on the side
    // thread while it saves ~0.1 s of first-launch cost — it only pays when parsing takes longer than that (inputs of tens of GB).
    warm_plan.go = !args.has("dump-alignment") && std::getenv("MSWB_WARM") != nullptr;
    warm_plan.ready = true;
  }
  warm_plan.cv.notify_all();
  // (declared after everything the plan refers to: on an early return the side threads are joined before those locals die)
  struct JoinBeforeLocalsDie { std::vector<std::thread> &t; ~JoinBeforeLocalsDie() { for (auto &x : t) if (x.joinable()) x.join(); } } join_early{warmups};

  // ---- pseudoalignments (src/mSWEEP.cpp:296-331) -------------------------------------------------------
  b200::ReadTable reads;
  try {
    log("  reading pseudoalignments");
    std::vector<std::string> paths;
    if (args.has("themisto")) paths = split(args.str("themisto", ""), ',');
    else paths = {args.str("themisto-1", ""), args.str("themisto-2", "")};
    reads = b200::read_themisto(paths, grouping.group_of_target.size(), args.str("themisto-mode", "intersection"), n_threads);
    log("  read alignments for " + std::to_string(reads.n_reads) + " reads");
  } catch (std::exception &e) {
    std::cerr << "Reading the pseudoalignments failed:\n  " << e.what() << "\nexiting\n";
    return 1;
  }
  const double t_parse = timer.lap();
  if (args.has("dump-alignment")) {   // debugging aid: the strand-merged alignment as CSR (u64 R, u64 T, row_ptr, targets)
    std::ofstream of(args.str("dump-alignment", ""), std::ios::binary);
    const uint64_t hdr[2] = {reads.n_reads, reads.n_targets};
    of.write((const char *)hdr, sizeof(hdr));
    of.write((const char *)reads.row_ptr.data(), (std::streamsize)(reads.row_ptr.size() * sizeof(uint64_t)));
    of.write((const char *)reads.targets.data(), (std::streamsize)(reads.targets.size() * sizeof(uint32_t)));
    return 0;
  }


  // One host thread per GPU.  Plain estimate: classes sharded over the GPUs (world = n_gpus, one NCCL
  // all-reduce per pass).  Bootstrap: every GPU holds the whole likelihood and takes replicates r % n_gpus.
  const int world = bootstrap ? 1 : n_gpus;
  std::vector<mswb_ctx *> group(n_gpus, nullptr);   // sharded mode: one NCCL clique created in one call
  if (world > 1) {
    for (auto &x : warmups) if (x.joinable()) x.join();
    std::vector<int> devs(n_gpus);
    for (int g = 0; g < n_gpus; ++g) devs[g] = g;
    if (mswb_ctx_create_group(n_gpus, devs.data(), group.data())) {
      std::cerr << "Initialising the GPUs failed:\n  " << mswb_last_error() << "\nexiting\n";
      return 1;
    }
  }

  // Sharded mode: every GPU builds only the classes of its own hash range (mswb_ec_build_partitioned).  The host routes
  // each aligned read by the reference's pattern hash (include/mSWEEP_alignment.hpp:150-155): rank r owns the r-th of
  // `world` equal ranges of the 64-bit hash space, so the ranks' tables, concatenated, are the global table in the
  // reference's order (ascending hash) and no class straddles two GPUs.  Unaligned reads belong to no class and stay
  // on the host (they only count in #num_reads).  part_ids[g][local row] = the read's id in the input.
  std::vector<b200::ReadTable> parts(world > 1 ? n_gpus : 0);
  std::vector<std::vector<uint32_t>> part_ids(world > 1 ? n_gpus : 0);
  double t_route = 0;
  if (world > 1) {
    Timer tr;
    const uint64_t R = reads.n_reads;
    std::vector<uint8_t> owner(R);
    std::vector<std::vector<uint64_t>> cnt(n_threads, std::vector<uint64_t>(2 * (size_t)n_gpus, 0));   // [thread][gpu]: reads, hits
#pragma omp parallel for num_threads(n_threads) schedule(static, 1)
    for (int tid = 0; tid < n_threads; ++tid) {      // one contiguous chunk of reads per (nominal) thread
      const int nth = n_threads;
      const uint64_t lo = R * tid / nth, hi = R * (tid + 1) / nth;
      for (uint64_t r = lo; r < hi; ++r) {
        const uint64_t a = reads.row_ptr[r], b = reads.row_ptr[r + 1];
        if (a == b) { owner[r] = 255; continue; }
        const uint64_t h = mswb_pattern_hash(reads.targets.data() + a, b - a);
        const int g = (int)(((unsigned __int128)h * (unsigned)n_gpus) >> 64);
        owner[r] = (uint8_t)g;
        cnt[tid][2 * g] += 1;
        cnt[tid][2 * g + 1] += b - a;
      }
    }
    // exclusive offsets per (thread, gpu) so that reads keep their input order inside every partition
    std::vector<std::vector<uint64_t>> off(n_threads, std::vector<uint64_t>(2 * (size_t)n_gpus, 0));
    for (int g = 0; g < n_gpus; ++g) {
      uint64_t nr = 0, nh = 0;
      for (int t = 0; t < n_threads; ++t) { off[t][2 * g] = nr; off[t][2 * g + 1] = nh; nr += cnt[t][2 * g]; nh += cnt[t][2 * g + 1]; }
      parts[g].n_reads = nr; parts[g].n_targets = reads.n_targets;
      parts[g].row_ptr.assign(nr + 1, 0);
      parts[g].targets.resize(nh);
      part_ids[g].resize(nr);
      parts[g].row_ptr[nr] = nh;
    }
#pragma omp parallel for num_threads(n_threads) schedule(static, 1)
    for (int tid = 0; tid < n_threads; ++tid) {
      const int nth = n_threads;
      std::vector<uint64_t> o = off[tid];
      const uint64_t lo = R * tid / nth, hi = R * (tid + 1) / nth;
      for (uint64_t r = lo; r < hi; ++r) {
        const int g = owner[r];
        if (g == 255) continue;
        const uint64_t a = reads.row_ptr[r], b = reads.row_ptr[r + 1];
        parts[g].row_ptr[o[2 * g]] = o[2 * g + 1];
        part_ids[g][o[2 * g]] = (uint32_t)r;
        std::copy(reads.targets.begin() + a, reads.targets.begin() + b, parts[g].targets.begin() + o[2 * g + 1]);
        o[2 * g] += 1;
        o[2 * g + 1] += b - a;
      }
    }
    t_route = tr.lap();
  }
  // A failure on one GPU must not leave the others waiting in a collective: the failing worker aborts every
  // communicator (mswb_ctx_abort), the peers' pending collectives end with an error, and main() reports the FIRST failure.
  std::atomic<int> first_failed{-1};

  std::vector<std::vector<std::vector<double>>> results_by_gpu(n_gpus);   // [gpu][0 = plain, 1.. = replicates][group]
  const bool want_probs = args.has("write-probs") || args.has("print-probs");
  const bool bin_reads = args.has("bin-reads");
  std::vector<std::vector<std::vector<uint32_t>>> bins_by_gpu(n_gpus);   // [gpu][kept group][read id]: bins of the GPU's class shard
  std::vector<std::string> probs_rows(n_gpus);   // formatted rows of each GPU's class shard, in class order
  std::vector<std::string> errors(n_gpus);
  std::vector<int> failed_stage(n_gpus, 0);   // 1 = EC/likelihood, 2 = estimation, 3 = bootstrap, 4 = binning
  std::vector<bool> mask;
  std::vector<uint64_t> ecs_by_gpu(n_gpus, 0), aligned_by_gpu(n_gpus, 0);
  uint64_t n_ecs = 0, n_aligned = 0, n_reads = 0;
  double t_ec = 0, t_lik = 0, t_vi = 0, t_boot = 0;
  b200::ViReport report;

  auto worker = [&](int gpu) {
    try {
      failed_stage[gpu] = 1;
      std::unique_ptr<b200::Context> ctx_holder(world > 1 ? new b200::Context(group[gpu], gpu, world) : new b200::Context(gpu));
      b200::Context &ctx = *ctx_holder;
      Timer tm;
      if (gpu == 0) log("Building equivalence classes");
      b200::Alignment aln(ctx, world > 1 ? parts[gpu] : reads, world > 1);
      ecs_by_gpu[gpu] = aln.n_ecs(); aligned_by_gpu[gpu] = aln.n_aligned();
      if (gpu == 0) { t_ec = tm.lap(); log("Computing the likelihood matrix"); }
      if (const char *inj = std::getenv("MSWB_TEST_FAIL_GPU"))   // fault injection for the abort path (tests only)
        if (std::atoi(inj) == gpu) throw std::runtime_error("injected failure on GPU " + std::to_string(gpu));
      // (a wide grouping that --min-hits prunes to <= 4096 groups takes the sparse form after all: the mask is global —
      //  all-reduced tallies — so every GPU takes the same decision)
      const bool try_sparse = sparse_if_pruned && min_hits > 0;
      std::unique_ptr<b200::Likelihood> ll_holder(new b200::Likelihood(ctx, aln, grouping.group_of_target, grouping.sizes, q, e_disp, min_hits, zi,
                                                                       try_sparse ? MSWB_STORE_SPARSE : storage));
      if (try_sparse && ll_holder->get_rows() > 4096) {
        ll_holder.reset();
        ll_holder.reset(new b200::Likelihood(ctx, aln, grouping.group_of_target, grouping.sizes, q, e_disp, min_hits, zi, storage));
      }
      b200::Likelihood &ll = *ll_holder;
      const std::vector<bool> my_mask = ll.groups_considered();
      if (gpu == 0) { mask = my_mask; t_lik = tm.lap(); }
      if (args.has("no-fit-model")) { failed_stage[gpu] = 0; return; }

      // prior counts (src/mSWEEP.cpp:391-398)
      std::vector<double> prior(ll.get_rows(), 1.0);
      if (args.has("alphas")) {
        std::vector<double> a;
        for (auto &x : split(args.str("alphas", ""), ',')) a.push_back(std::stod(x));
        if (a.size() != ll.get_rows()) throw std::runtime_error("--alphas must have the same number of values as there are groups.");
        prior = a;
      }
      failed_stage[gpu] = 2;
      if (gpu == 0) log("Estimating relative abundances");
      std::vector<std::vector<double>> res;
      b200::ViReport rep;
      res.push_back(b200::rcg_optl(ctx, ll, nullptr, prior, vi, (verbose && gpu == 0) ? &std::cerr : nullptr, &rep));
      if (gpu == 0) { report = rep; t_vi = tm.lap(); }
      if (want_probs && (!bootstrap || gpu == 0)) {
        // src/Sample.cpp:63-85 / 154-186: one row per class, exp(log-posterior) per group, pruned groups as 0 at the
        // end.  The K x N matrix never sits on the host: tiles of classes are pulled and formatted as they come.
        const size_t n_zero = (size_t)std::count(my_mask.begin(), my_mask.end(), false);
        std::ostringstream os;
        const uint64_t tile = 4096, N = ll.get_cols(), K = ll.get_rows();
        for (uint64_t b = 0; b < N; b += tile) {
          const uint64_t e = std::min(N, b + tile), n = e - b;
          const std::vector<double> g = ll.posteriors(b, e);
          for (uint64_t j = 0; j < n; ++j) {
            os << (ll.ec_begin() + b + j) << '\t';
            for (uint64_t k = 0; k < K; ++k) os << std::exp(g[k * n + j]) << (k + 1 < K + n_zero ? '\t' : '\n');
            for (size_t z = 0; z < n_zero; ++z) os << (double)0.0 << (z + 1 < n_zero ? '\t' : '\n');
          }
        }
        probs_rows[gpu] = os.str();
      }
      if (bin_reads && (!bootstrap || gpu == 0)) {
        // src/mSWEEP.cpp:437-456: targets default to every estimated group; --min-abundance filters them
        // (mGEMS::FilterTargetGroups); a class joins a target's bin when its posterior reaches the group's abundance.
        failed_stage[gpu] = 4;
        std::vector<std::string> est_names;
        for (size_t g = 0; g < grouping.names.size(); ++g) if (my_mask[g]) est_names.push_back(grouping.names[g]);
        std::set<std::string> targets(est_names.begin(), est_names.end());
        if (args.has("target-groups")) {
          targets.clear();
          for (auto &name : split(args.str("target-groups", ""), ',')) {
            if (std::find(est_names.begin(), est_names.end(), name) == est_names.end())
              throw std::runtime_error("Target group " + name + " is not among the estimated groups.");
            targets.insert(name);
          }
        }
        const double min_abundance = args.num<double>("min-abundance", 0.0);
        std::vector<double> thr(ll.get_rows(), std::numeric_limits<double>::infinity());
        for (size_t k = 0; k < est_names.size(); ++k)
          if (targets.count(est_names[k]) && !(args.has("min-abundance") && res[0][k] < min_abundance)) thr[k] = std::log(res[0][k]);
        bins_by_gpu[gpu] = ll.assign(aln, thr);
        if (world > 1)   // the partition's rows are local: back to the read ids of the input
          for (auto &bin : bins_by_gpu[gpu]) for (auto &id : bin) id = part_ids[gpu][id];
        failed_stage[gpu] = 2;
      }
      if (bootstrap) {
        failed_stage[gpu] = 3;
        if (gpu == 0) log("Running estimation with " + std::to_string(iters) + " bootstrap iterations");
        auto reps = ll.bootstrap(prior, vi, iters, bootstrap_count, seed, gpu, n_gpus, rng_mode);
        for (auto &r : reps) res.push_back(std::move(r));
        if (gpu == 0) t_boot = tm.lap();
      }
      results_by_gpu[gpu] = std::move(res);
      failed_stage[gpu] = 0;
    } catch (const std::exception &e) {
      errors[gpu] = e.what();
      int none = -1;
      if (first_failed.compare_exchange_strong(none, gpu) && world > 1)
        for (int g = 0; g < n_gpus; ++g) mswb_ctx_abort(group[g]);
    }
  };
  for (auto &x : warmups) if (x.joinable()) x.join();
  const double t_warm = timer.lap();
  {
    std::vector<std::thread> threads;
    for (int g = 1; g < n_gpus; ++g) threads.emplace_back(worker, g);
    worker(0);
    for (auto &t : threads) t.join();
  }
  for (int i = 0; i < n_gpus; ++i) {
    const int g = first_failed.load() >= 0 ? (first_failed.load() + i) % n_gpus : i;   // the original failure first
    if (!failed_stage[g]) continue;
    const char *what = failed_stage[g] == 1 ? "Building the log-likelihood array failed:\n  "
                     : failed_stage[g] == 2 ? "Estimating relative abundances failed:\n  "
                     : failed_stage[g] == 4 ? "Binning the reads failed:\n  " : "Bootstrap iteration failed:\n  ";
    std::cerr << what << errors[g] << "\nexiting\n";
    return 1;
  }
  n_reads = reads.n_reads;
  if (world > 1) for (int g = 0; g < n_gpus; ++g) { n_ecs += ecs_by_gpu[g]; n_aligned += aligned_by_gpu[g]; }
  else { n_ecs = ecs_by_gpu[0]; n_aligned = aligned_by_gpu[0]; }
  log("  found " + std::to_string(n_ecs) + " unique alignments");
  if (args.has("no-fit-model")) { log("Skipping relative abundance estimation (--no-fit-model toggled)"); return 0; }
  if (min_hits > 0)
    std::cerr << "WARNING: --min-hits > 0 is an experimental option that has not been thoroughly tested and is subject to change.\n" << std::endl;

  // merge the replicates computed on the different GPUs
  std::vector<std::vector<double>> results = results_by_gpu[0];
  for (uint64_t r = 0; r < iters; ++r) results[r + 1] = results_by_gpu[r % n_gpus][r + 1];

  // names of estimated vs pruned groups (src/mSWEEP.cpp:425-435)
  std::vector<std::string> estimated, zero;
  for (size_t g = 0; g < grouping.names.size(); ++g) (mask[g] ? estimated : zero).push_back(grouping.names[g]);

  if (args.has("run-rate"))
    std::cerr << "WARNING: --run-rate is an experimental option that has not been thoroughly tested and is subject to change.\n" << std::endl;

  if (bin_reads) {
    // one file per target group in the directory -o points to (src/OutfileDesignator.cpp:80-94), one read per line.
    // Read ids are written 1-based, the numbering mGEMS extract counts fastq records in (mGEMS is off-tree: assumption).
    std::string dir = ".";
    const std::string o = args.str("o", "");
    if (o.find('/') != std::string::npos) dir = o.substr(0, o.rfind('/'));
    for (size_t k = 0; k < estimated.size(); ++k) {
      std::vector<uint32_t> bin;
      for (int g = 0; g < (bootstrap ? 1 : n_gpus); ++g) bin.insert(bin.end(), bins_by_gpu[g][k].begin(), bins_by_gpu[g][k].end());
      if (n_gpus > 1 && !bootstrap) std::sort(bin.begin(), bin.end());
      bool wanted = !args.has("target-groups");
      if (!wanted) { const auto t = split(args.str("target-groups", ""), ','); wanted = std::find(t.begin(), t.end(), estimated[k]) != t.end(); }
      if (wanted && args.has("min-abundance") && results[0][k] < args.num<double>("min-abundance", 0.0)) wanted = false;
      if (!wanted) continue;
      std::ofstream of(dir + '/' + estimated[k] + ".bin");
      if (!of.good()) { std::cerr << "Writing the bin for target group " << estimated[k] << " failed:\n  Can't write to bin file.\nexiting\n"; return 1; }
      std::string text;
      for (uint32_t r : bin) { text += std::to_string((uint64_t)r + 1); text += '\n'; }
      of << text;
    }
  }

  if (want_probs) {
    try {
      auto write_probs = [&](std::ostream &of) {
        if (!of.good()) throw std::runtime_error("Can't write to probs file.");
        of << "ec_id" << '\t';
        const size_t n_rows = estimated.size() + zero.size();
        for (size_t i = 0; i < n_rows; ++i)
          of << (i < estimated.size() ? estimated[i] : zero[i - estimated.size()]) << (i + 1 < n_rows ? '\t' : '\n');
        for (int g = 0; g < (bootstrap ? 1 : n_gpus); ++g) of << probs_rows[g];
        of << std::endl;
      };
      if (args.has("print-probs")) write_probs(std::cout);
      if (args.has("write-probs")) { std::ofstream of(args.str("o", "") + "_probs.tsv"); write_probs(of); }   // src/OutfileDesignator.cpp:96-102
    } catch (std::exception &e) {
      std::cerr << "Writing the probabilities failed:\n  " << e.what() << "\nexiting\n";
      return 1;
    }
  }

  try {
    const std::string o = args.str("o", "");
    std::ofstream file;
    if (!o.empty()) file.open(o + "_abundances.txt");                 // src/OutfileDesignator.cpp:104-114
    std::ostream &of = o.empty() ? std::cout : file;
    if (args.has("run-rate")) {
      // src/mSWEEP.cpp:524-548: mean_theta, RATE and KLD per group; the bootstrap columns are not written in this mode
      if (!of.good()) throw std::runtime_error("Can't write to abundances file.");
      const auto kld = b200::dirichlet_kld(results[0], (double)n_aligned);
      of << "#mSWEEP_version:" << '\t' << MSWEEP_BUILD_VERSION << '\n';
      of << "#num_reads:" << '\t' << n_reads << '\n';
      of << "#num_aligned:" << '\t' << n_aligned << '\n';
      of << "#c_id" << '\t' << "mean_theta" << '\t' << "RATE" << '\t' << "KLD" << '\n';
      for (size_t i = 0; i < estimated.size(); ++i)
        of << estimated[i] << '\t' << results[0][i] << '\t' << kld.second[i] << '\t' << std::exp(kld.first[i]) << '\n';
      for (size_t i = 0; i < zero.size(); ++i) of << zero[i] << '\t' << (double)0.0 << '\t' << (double)0.0 << '\t' << (double)0.0 << '\n';
      of.flush();
    } else {
      write_abundances(of, n_reads, n_aligned, estimated, zero, results, iters);
    }
  } catch (std::exception &e) {
    std::cerr << "Writing the relative abundances failed:\n  " << e.what() << "\nexiting\n";
    return 1;
  }
  const double t_gpu_total = timer.lap();
  if (timings)
    std::cerr << "{\"grouping_s\": " << t_grouping << ", \"parse_s\": " << t_parse << ", \"cuda_init_wait_s\": " << t_warm
              << ", \"route_s\": " << t_route << ", \"gpu_stage_total_s\": " << t_gpu_total << ", \"ec_build_s\": " << t_ec
              << ", \"likelihood_s\": " << t_lik << ", \"optimiser_s\": " << t_vi << ", \"bootstrap_s\": " << t_boot
              << ", \"write_s\": " << timer.lap() << ", \"n_ecs\": " << n_ecs << ", \"iters\": " << report.iters
              << ", \"bound\": " << report.bound << ", \"converged\": " << (report.converged ? "true" : "false") << "}" << std::endl;
  return 0;
}
