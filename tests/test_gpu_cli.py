"""GPU: the mSWEEP_b200 binary end to end (files in, <prefix>_abundances.txt out) against the oracle's
command-line front end on the same files: the reference's output format byte for byte (modulo the
version line), abundances within the printed precision."""
import os
import subprocess

import numpy as np
import pytest

from msweep_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "msweep_b200", "bin", "mSWEEP_b200")
ORACLE = os.path.join(ROOT, "oracle", "msweep_oracle")


@pytest.fixture(scope="module")
def data(tmp_path_factory):
    d = tmp_path_factory.mktemp("cli")
    wl = synth.generate(20000, 600, 30, n_present=4, n_templates=200, p_noise=0.02, seed=31)
    paths = synth.write_themisto(str(d / "aln"), wl, paired=True, shuffle_frac=0.05)
    g = str(d / "grouping.txt")
    synth.write_grouping(g, wl)
    return d, wl, paths, g


def parse_abundances(path):
    head, rows = [], []
    for line in open(path).read().splitlines():
        (head if line.startswith("#") else rows).append(line)
    names = [r.split("\t")[0] for r in rows]
    vals = np.array([[float(x) for x in r.split("\t")[1:]] for r in rows])
    return head, names, vals


def run_both(d, paths, g, extra, tag, oracle_algo="rcgcpu"):
    ours, ref = str(d / f"ours_{tag}"), str(d / f"ref_{tag}")
    r = subprocess.run([CLI, "--themisto-1", paths[0], "--themisto-2", paths[1], "-i", g, "-o", ours, "-t", "4", *extra],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    ex = [x for x in extra if x not in ("rcgb200", "emb200", "sparse", "dense")]
    for flag in ("--algorithm", "--storage"):
        if flag in ex:
            ex.remove(flag)
    r2 = subprocess.run([ORACLE, "--themisto-1", paths[0], "--themisto-2", paths[1], "-i", g, "-o", ref, "-t", "4",
                         "--algorithm", oracle_algo, *ex], capture_output=True, text=True)
    assert r2.returncode == 0, r2.stderr
    return parse_abundances(ours + "_abundances.txt"), parse_abundances(ref + "_abundances.txt"), r.stderr


def test_plain_estimate(data):
    d, wl, paths, g = data
    (h, names, vals), (h2, names2, vals2), _ = run_both(d, paths, g, [], "plain")
    assert h[0].startswith("#mSWEEP_version:\t")
    assert h[1:] == h2[1:]                                  # num_reads, num_aligned, column header: identical bytes
    assert h[1] == f"#num_reads:\t{wl.n_reads}" and h[-1] == "#c_id\tmean_theta"
    assert names == names2 == wl.group_names
    assert np.max(np.abs(vals - vals2)) < 2e-6              # 6 significant digits are printed
    assert abs(vals.sum() - 1.0) < 1e-5


def test_em_and_float_precision(data):
    d, wl, paths, g = data
    (h, names, vals), (h2, _, vals2), _ = run_both(d, paths, g, ["--algorithm", "emb200"], "em", oracle_algo="emgpu")
    assert h[1:] == h2[1:] and np.max(np.abs(vals - vals2)) < 2e-6
    (h, names, vals3), _, _ = run_both(d, paths, g, ["--algorithm", "emb200", "--emprecision", "float", "--tol", "1e-5"], "emf",
                                       oracle_algo="emgpu")
    assert np.max(np.abs(vals3 - vals2)) < 1e-3             # float EM stops elsewhere on the plateau (docs/gpubenchmarks.md:27)


def test_storage_flag(data):
    """--storage auto (the default) = the lossless sparse form for fp64 likelihoods; dense and sparse give the reference's
    numbers for both optimisers, probabilities included; sparse cannot be combined with --emprecision float."""
    d, wl, paths, g = data
    for algo, oracle_algo in (("emb200", "emgpu"), ("rcgb200", "rcgcpu")):
        for store in ("sparse", "dense"):
            (h, names, vals), (h2, _, vals2), _ = run_both(d, paths, g, ["--algorithm", algo, "--storage", store], "st_" + algo + store,
                                                          oracle_algo=oracle_algo)
            assert h[1:] == h2[1:] and np.max(np.abs(vals - vals2)) < 2e-6
    outs = {}
    for store in ("sparse", "dense"):
        r = subprocess.run([CLI, "--themisto", ",".join(paths), "-i", g, "--storage", store, "--print-probs"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        outs[store] = r.stdout
    rows = {k: [l.split("\t") for l in v.splitlines() if l and not l.startswith("#") and not l.startswith("ec_id") and l[0].isdigit()] for k, v in outs.items()}
    assert len(rows["sparse"]) == len(rows["dense"]) > 0
    for a, b in zip(rows["sparse"], rows["dense"]):
        assert a[0] == b[0] and np.max(np.abs(np.array(a[1:], float) - np.array(b[1:], float))) < 2e-6
    r = subprocess.run([CLI, "--themisto", ",".join(paths), "-i", g, "--storage", "sparse", "--algorithm", "emb200", "--emprecision", "float"],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "fp64 form" in r.stderr
    r = subprocess.run([CLI, "--themisto", ",".join(paths), "-i", g, "--storage", "compact"], capture_output=True, text=True)
    assert r.returncode == 1 and "Unknown --storage" in r.stderr


def test_min_hits_orders_pruned_groups_last(data):
    d, wl, paths, g = data
    (h, names, vals), (h2, names2, vals2), err = run_both(d, paths, g, ["--min-hits", "200"], "mh")
    assert "WARNING: --min-hits > 0 is an experimental option" in err
    assert names == names2 and sorted(names) == sorted(wl.group_names)
    assert np.max(np.abs(vals - vals2)) < 2e-6
    n_zero = int((vals[:, 0] == 0).sum())
    assert n_zero > 0 and np.all(vals[-n_zero:, 0] == 0) and np.all(vals[:-n_zero, 0] > 0)   # src/PlainSample.cpp:56-66


def test_wide_grouping_pruned_by_min_hits_takes_the_sparse_form(tmp_path):
    """Config 4's shape in small: more groups than the sparse sweeps hold in shared memory (> 4096), nearly all of them
    empty.  --storage auto builds the sparse form once --min-hits has pruned the grouping; the abundances equal the dense
    form's and the oracle's, and without --min-hits the same input stays dense (and still runs)."""
    wl = synth.generate(6000, 10000, 5000, n_present=4, n_templates=100, p_noise=0.0, seed=41)
    paths = synth.write_themisto(str(tmp_path / "aln"), wl, paired=True)
    g = str(tmp_path / "grouping.txt")
    synth.write_grouping(g, wl)
    (h, names, v_auto), (h2, names2, v_ref), _ = run_both(tmp_path, paths, g, ["--min-hits", "1"], "wide_auto")
    (_, names_d, v_dense), _, _ = run_both(tmp_path, paths, g, ["--min-hits", "1", "--storage", "dense"], "wide_dense")
    assert names == names2 == names_d and h[1:] == h2[1:]
    assert np.max(np.abs(v_auto - v_ref)) < 2e-6 and np.max(np.abs(v_auto - v_dense)) < 2e-6
    r = subprocess.run([CLI, "--themisto-1", paths[0], "--themisto-2", paths[1], "-i", g, "-o", str(tmp_path / "wide_nomh"), "-t", "4"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_bootstrap_output(data):
    d, wl, paths, g = data
    (h, names, vals), (h2, names2, vals2), _ = run_both(d, paths, g, ["--iters", "3", "--seed", "11"], "boot")
    assert h[1:] == h2[1:]
    assert h[3] == "#bootstrap_iters:\t3" and h[4] == "#c_id\tmean_theta\tbootstrap_mean_thetas"
    assert vals.shape == (len(wl.group_names), 4)
    assert np.max(np.abs(vals - vals2)) < 2e-6              # same mt19937_64 stream -> same resampled counts -> same estimates


def test_options_reach_the_model(data):
    d, wl, paths, g = data
    extra = ["-q", "0.5", "-e", "0.02", "--zero-inflation", "0.02", "--alphas", ",".join(["0.7"] * wl.n_groups), "--tol", "1e-8",
             "--themisto-mode", "union"]
    (h, names, vals), (h2, names2, vals2), _ = run_both(d, paths, g, extra, "opts")
    assert h[1:] == h2[1:] and np.max(np.abs(vals - vals2)) < 2e-6


def test_stdout_when_no_prefix_and_timings(data):
    d, wl, paths, g = data
    r = subprocess.run([CLI, "--themisto", ",".join(paths), "-i", g, "--print-timings"], capture_output=True, text=True)
    assert r.returncode == 0
    assert r.stdout.startswith("#mSWEEP_version:\t") and r.stdout.count("\n") == 4 + wl.n_groups
    assert '"optimiser_s"' in r.stderr


def test_alphas_length_is_checked(data):
    d, wl, paths, g = data
    r = subprocess.run([CLI, "--themisto", ",".join(paths), "-i", g, "--alphas", "1,1"], capture_output=True, text=True)
    assert r.returncode == 1 and "--alphas must have the same number of values as there are groups." in r.stderr


def test_write_probs(data):
    """src/Sample.cpp:63-85, 154-186: <prefix>_probs.tsv, one row per class, pruned groups as trailing zeros."""
    d, wl, paths, g = data
    for extra, tag in (([], "p0"), (["--min-hits", "200"], "p1")):
        run_both(d, paths, g, ["--write-probs", *extra], tag)
        ours = open(d / f"ours_{tag}_probs.tsv").read().split("\n")
        ref = open(d / f"ref_{tag}_probs.tsv").read().split("\n")
        assert ours[0] == ref[0] and ours[0].startswith("ec_id\t")
        assert len(ours) == len(ref) and ours[-1] == "" and ours[-2] == ""          # trailing std::endl
        a = np.array([[float(x) for x in l.split("\t")] for l in ours[1:-2]])
        b = np.array([[float(x) for x in l.split("\t")] for l in ref[1:-2]])
        assert a.shape == b.shape and np.array_equal(a[:, 0], np.arange(len(a)))
        assert np.max(np.abs(a - b)) < 2e-6
        assert np.max(np.abs(a[:, 1:].sum(axis=1) - 1.0)) < 1e-4


def test_run_rate_columns(data):
    """--run-rate (src/Sample.cpp:99-151, src/mSWEEP.cpp:524-548): mean_theta, RATE and KLD per group."""
    d, wl, paths, g = data
    (h, names, vals), (h2, names2, vals2), err = run_both(d, paths, g, ["--run-rate"], "rate")
    assert "WARNING: --run-rate is an experimental option" in err
    assert h[-1] == h2[-1] == "#c_id\tmean_theta\tRATE\tKLD"
    assert names == names2 and vals.shape == vals2.shape == (len(names), 3)
    assert np.max(np.abs(vals[:, 0] - vals2[:, 0])) < 2e-6
    assert np.allclose(vals[:, 1], vals2[:, 1], rtol=1e-4, atol=1e-9)      # RATE
    assert np.allclose(vals[:, 2], vals2[:, 2], rtol=1e-4, atol=1e-12)     # KLD
    assert abs(vals[:, 1].sum() - 1.0) < 1e-4


def test_bin_reads_files(data):
    """--bin-reads: one <group>.bin per target group next to the -o prefix, the reads of every class whose posterior
    reaches the group's abundance."""
    d, wl, paths, g = data
    dirs = []
    for who in ("ours", "ref"):
        (d / f"bins_{who}").mkdir()
        dirs.append(d / f"bins_{who}")
    common = ["--themisto-1", paths[0], "--themisto-2", paths[1], "-i", g, "-t", "4", "--bin-reads", "--min-abundance", "0.01"]
    r = subprocess.run([CLI, *common, "-o", str(dirs[0] / "x")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r2 = subprocess.run([ORACLE, *common, "-o", str(dirs[1] / "x")], capture_output=True, text=True)
    assert r2.returncode == 0, r2.stderr
    ours = sorted(f for f in os.listdir(dirs[0]) if f.endswith(".bin"))
    ref = sorted(f for f in os.listdir(dirs[1]) if f.endswith(".bin"))
    assert ours == ref and 1 <= len(ours) < len(wl.group_names)       # --min-abundance drops the absent groups
    total = moved = 0
    for f in ours:
        a = np.loadtxt(dirs[0] / f, dtype=np.int64, ndmin=1)
        b = np.loadtxt(dirs[1] / f, dtype=np.int64, ndmin=1)
        assert np.all(np.diff(a) > 0) and a.min() >= 1 and a.max() <= wl.n_reads
        total += b.size
        moved += np.setxor1d(a, b).size
    assert total > 0 and moved <= 1e-3 * total
    # --target-groups restricts the bins; an unknown group fails like the reference's binning stage does
    (d / "bins_one").mkdir()
    one = ours[0][:-4]
    r = subprocess.run([CLI, *common[:-2], "--target-groups", one, "-o", str(d / "bins_one" / "x")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert sorted(f for f in os.listdir(d / "bins_one") if f.endswith(".bin")) == [one + ".bin"]
    r = subprocess.run([CLI, *common[:-2], "--target-groups", "no_such_group", "-o", str(d / "bins_one" / "y")], capture_output=True, text=True)
    assert r.returncode == 1 and r.stderr.startswith("Binning the reads failed:")


def test_rcg_optl_dense_has_the_rcgpar_argument_list(oracle, tmp_path):
    """b200::rcg_optl_dense(ctx, logl K x N, log_times_observed, alpha0, tol, max_iters, ostream&) -> K x N log-posteriors: what
    rcgpar::rcg_optl_omp / em_torch return at src/mSWEEP.cpp:194-202, so that an unmodified call site can bind."""
    rng = np.random.default_rng(3)
    K, N = 9, 400
    logl = rng.normal(-6.0, 2.0, size=(K, N))
    logl[rng.integers(0, K, size=N), np.arange(N)] = -0.4
    lc = np.log(rng.integers(1, 30, size=N).astype(np.float64))
    alpha0 = rng.uniform(0.5, 2.0, size=K)
    np.concatenate([logl.ravel(), lc, alpha0]).tofile(tmp_path / "in.bin")
    src = tmp_path / "dense.cpp"
    src.write_text('''#include "msweep_b200.hpp"
#include <cstdio>
#include <fstream>
#include <iostream>
int main(int argc, char **argv) {
  const uint32_t K = std::atoi(argv[2]); const uint64_t N = std::atoll(argv[3]); const int algo = std::atoi(argv[4]);
  std::vector<double> buf((size_t)K * N + N + K);
  std::ifstream(argv[1], std::ios::binary).read((char *)buf.data(), buf.size() * sizeof(double));
  std::vector<double> logl(buf.begin(), buf.begin() + (size_t)K * N), lc(buf.begin() + (size_t)K * N, buf.begin() + (size_t)K * N + N),
      a0(buf.end() - K, buf.end()), theta;
  try {
    b200::Context ctx(0);
    std::ofstream quiet;                                   // never opened: the reference's non-verbose log (src/mSWEEP.cpp:190)
    std::vector<double> gamma = b200::rcg_optl_dense(ctx, logl, K, N, lc, a0, 1e-6, 5000, quiet, algo, MSWB_STORE_F64, &theta);
    fwrite(gamma.data(), sizeof(double), gamma.size(), stdout);
    fwrite(theta.data(), sizeof(double), theta.size(), stdout);
    try { b200::rcg_optl_dense(ctx, logl, K, N + 1, lc, a0, 1e-6, 10, quiet); return 3; } catch (const std::runtime_error &) {}
  } catch (const std::exception &e) { std::cerr << e.what() << std::endl; return 1; }
  return 0;
}
''')
    exe = tmp_path / "dense"
    lib = os.path.join(ROOT, "msweep_b200", "lib")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "msweep_b200", "host"),
                    str(src), "-o", str(exe), "-L", lib, "-lmsweep_b200", f"-Wl,-rpath,{lib}"], check=True)
    for algo, name in ((0, "rcg"), (1, "em")):
        out = subprocess.run([str(exe), str(tmp_path / "in.bin"), str(K), str(N), str(algo)], capture_output=True, check=True).stdout
        got = np.frombuffer(out, np.float64)
        gamma, theta = got[:K * N].reshape(K, N), got[K * N:]
        ref = oracle.vi_run(name, logl, lc, alpha0=alpha0, want_gamma=True)
        assert np.max(np.abs(theta - ref.theta)) < 1e-6
        assert np.max(np.abs(np.exp(gamma) - np.exp(ref.gamma))) < 1e-6
        assert np.max(np.abs(np.exp(gamma).sum(axis=0) - 1.0)) < 1e-12


def test_bootstrap_count_is_honoured(data):
    """--bootstrap-count: the flag's value is the number of pseudoalignments resampled per replicate (help text, src/mSWEEP.cpp:141).
    The reference's ConstructSample passes --iters in its place unless --bin-reads is given (src/Sample.cpp:38-39, SURVEY §9: a quirk);
    this binary and the oracle CLI both use the flag's value — an intentional deviation recorded in DESIGN.md §3.  Fewer draws give noisier replicates."""
    d, wl, paths, g = data
    (h, names, vals), (h2, _, vals2), _ = run_both(d, paths, g, ["--iters", "4", "--seed", "5", "--bootstrap-count", "300"], "bc")
    assert h[1:] == h2[1:] and np.max(np.abs(vals - vals2)) < 2e-6
    (_, _, full), _, _ = run_both(d, paths, g, ["--iters", "4", "--seed", "5"], "bcfull")
    top = int(np.argmax(full[:, 0]))
    assert np.std(vals[top, 1:]) > 3 * np.std(full[top, 1:])          # 300 draws against all aligned reads
