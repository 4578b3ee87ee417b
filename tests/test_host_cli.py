"""CPU: the C++17 host driver (msweep_b200/bin/mSWEEP_b200) up to the point where it needs the GPU:
argument handling with the reference's messages and exit codes, the -i reader, and the multi-threaded
Themisto parser against the oracle's line-by-line restatement of include/mSWEEP_alignment.hpp:54-135."""
import os
import subprocess

import numpy as np
import pytest

from msweep_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "msweep_b200", "bin", "mSWEEP_b200")


def run(*args):
    return subprocess.run([CLI, *args], capture_output=True, text=True)


def dump(tmp_path, paths, grouping, mode="intersection", threads=4):
    out = str(tmp_path / "aln.bin")
    r = run("--themisto", ",".join(paths), "-i", grouping, "--themisto-mode", mode, "-t", str(threads), "--dump-alignment", out)
    assert r.returncode == 0, r.stderr
    raw = open(out, "rb").read()
    R, T = np.frombuffer(raw[:16], np.uint64)
    rp = np.frombuffer(raw[16:16 + 8 * (int(R) + 1)], np.uint64)
    tg = np.frombuffer(raw[16 + 8 * (int(R) + 1):], np.uint32)
    return int(R), int(T), rp, tg


def test_binary_exists():
    assert os.path.exists(CLI), "build with python -m msweep_b200._build"


def test_version_help_cite():
    assert run("--version").returncode == 0 and "mSWEEP-" in run("--version").stderr
    assert "Usage: mSWEEP_b200" in run("--help").stderr
    assert "Wellcome Open Res" in run("--cite").stderr


@pytest.mark.parametrize("argv,msg", [
    (["--themisto", "x"], "Required argument -i"),
    (["-i", "g"], "No pseudoalignment files"),
    (["-i", "g", "--themisto-1", "a"], "must be given together"),
    (["-i", "g", "--themisto", "a", "--bogus", "1"], "Unknown argument"),
    (["-i", "g", "--themisto", "a", "--write-likelihood"], "outside the scope"),
    (["-i", "g", "--themisto", "a", "-o", "/nonexistent_dir_xyz/out"], "does not seem to exist"),
])
def test_argument_errors(argv, msg):
    r = run(*argv)
    assert r.returncode == 1
    assert r.stderr.startswith("Error in parsing arguments:\n  ") and r.stderr.endswith("\nexiting\n")   # src/mSWEEP.cpp:240-245
    assert msg in r.stderr


def test_algorithm_is_validated(tmp_path):
    g = tmp_path / "g.txt"
    g.write_text("a\nb\n")
    for algo, msg in (("rcgcpu", "no CPU path"), ("foo", "Unknown --algorithm")):
        r = run("-i", str(g), "--themisto", "x", "--algorithm", algo)
        assert r.returncode == 1 and msg in r.stderr


def test_grouping_errors(tmp_path):
    r = run("-i", str(tmp_path / "missing.txt"), "--themisto", "x")
    assert r.returncode == 1 and r.stderr == "Reading group indicators failed:\n  Could not read cluster indicators.\nexiting\n"
    empty = tmp_path / "empty.txt"
    empty.write_text("")
    r = run("-i", str(empty), "--themisto", "x")
    assert r.returncode == 1 and "The grouping contains 0 reference sequences" in r.stderr


@pytest.mark.parametrize("mode", ["intersection", "union"])
@pytest.mark.parametrize("threads", [1, 5])
def test_parser_matches_oracle(oracle, tmp_path, mode, threads):
    wl = synth.generate(3000, 90, 6, n_present=3, n_templates=40, p_noise=0.05, seed=21)
    paths = synth.write_themisto(str(tmp_path / "aln"), wl, paired=True, shuffle_frac=0.2)
    gpath = str(tmp_path / "g.txt")
    synth.write_grouping(gpath, wl)
    R, T, rp, tg = dump(tmp_path, paths, gpath, mode, threads)
    n, rp_o, tg_o = oracle.read_themisto(paths, wl.n_targets, mode)
    assert (R, T) == (n, wl.n_targets)
    assert np.array_equal(rp, rp_o) and np.array_equal(tg, tg_o)
    if mode == "intersection":
        assert np.array_equal(rp, wl.row_ptr) and np.array_equal(tg, wl.targets)


def test_parser_quirks_match_reference_semantics(oracle, tmp_path):
    """Duplicate ids accumulate, unsorted and repeated targets collapse to a bit set, a target id >= T
    lands in a later read's row (flat bit index, mSWEEP_alignment.hpp:64), rows past the line count are
    ignored, a trailing space and CRLF are tolerated."""
    g = tmp_path / "g.txt"
    g.write_text("".join(f"g{i % 3}\n" for i in range(10)))
    a = tmp_path / "a.aln"
    a.write_text("0 5 3 3 1\n2 7 \n1\n0 9\n3 12\r\n7 1\n")
    R, T, rp, tg = dump(tmp_path, [str(a)], str(g))
    n, rp_o, tg_o = oracle.read_themisto([str(a)], 10, "intersection")
    assert R == n == 6
    assert np.array_equal(rp, rp_o) and np.array_equal(tg, tg_o)
    rows = [list(tg[int(rp[i]):int(rp[i + 1])]) for i in range(R)]
    assert rows[0] == [1, 3, 5, 9] and rows[2] == [7] and rows[1] == [] and rows[4] == [2] and rows[3] == []


def test_parser_fast_and_slow_paths_agree(oracle, tmp_path):
    """Well-formed lines take the tokeniser's fast path; a sign, a tab, junk glued to a number or a CR rewind the line into
    the std::stoul-compatible path.  Mixed at random (with target ids >= T, which spill into later rows) against the oracle."""
    rng = np.random.default_rng(12)
    T, R = 40, 3000
    g = tmp_path / "g.txt"
    g.write_text("".join(f"g{i % 4}\n" for i in range(T)))
    lines = []
    for r in rng.permutation(R):
        toks = [str(int(t)) for t in rng.integers(0, T + 25, size=rng.integers(0, 9))]
        if rng.random() < 0.3:
            for i in range(len(toks)):
                kind = rng.integers(0, 5)
                toks[i] = {0: "+" + toks[i], 1: "\t" + toks[i], 2: toks[i] + "x7", 3: toks[i] + ".5", 4: toks[i]}[int(kind)]
        line = " ".join([str(int(r))] + toks)
        trailing = rng.random() < 0.1
        if trailing:
            line += " "
        # (a CR after a trailing space would be a token of its own, which std::stoul rejects: "File format not supported")
        lines.append(line + ("\r\n" if not trailing and rng.random() < 0.05 else "\n"))
    a = tmp_path / "a.aln"
    a.write_text("".join(lines))
    for threads in (1, 5):
        Rn, Tn, rp, tg = dump(tmp_path, [str(a)], str(g), threads=threads)
        n, rp_o, tg_o = oracle.read_themisto([str(a)], T, "intersection")
        assert Rn == n == R
        assert np.array_equal(rp, rp_o) and np.array_equal(tg, tg_o)


@pytest.mark.parametrize("content,line", [("0 1\nx 2\n", 2), ("0 1\n1  2\n", 2), ("0 1\n\n", 2), ("0 a\n", 1)])
def test_parser_errors_name_the_line(tmp_path, content, line):
    g = tmp_path / "g.txt"
    g.write_text("a\nb\nc\n")
    a = tmp_path / "a.aln"
    a.write_text(content)
    r = run("--themisto", str(a), "-i", str(g), "-t", "3", "--dump-alignment", str(tmp_path / "o.bin"))
    assert r.returncode == 1
    assert r.stderr.startswith("Reading the pseudoalignments failed:\n  File format not supported on line %d with content: " % line)


def test_bad_merge_mode_and_compact_format(tmp_path):
    g = tmp_path / "g.txt"
    g.write_text("a\nb\nc\n")
    a = tmp_path / "a.aln"
    a.write_text("0 1\n1 2\n")
    r = run("--themisto", f"{a},{a}", "-i", str(g), "--themisto-mode", "unpaired", "--dump-alignment", str(tmp_path / "o.bin"))
    assert r.returncode == 1 and "Unrecognized option `unpaired` for --themisto-mode" in r.stderr
    c = tmp_path / "c.aln"
    c.write_text("n_reads:2,n_refs:3\n")
    r = run("--themisto", str(c), "-i", str(g), "--dump-alignment", str(tmp_path / "o.bin"))
    assert r.returncode == 1 and "compact" in r.stderr


def test_no_gpu_fails_loudly(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    g = tmp_path / "g.txt"
    g.write_text("a\nb\nc\n")
    a = tmp_path / "a.aln"
    a.write_text("0 1\n1 2\n")
    r = run("--themisto", str(a), "-i", str(g))
    assert r.returncode == 1 and "no CUDA device available" in r.stderr and r.stdout == ""


def test_gzip_input_is_read_transparently(oracle, tmp_path):
    """The reference opens its inputs through cxxio (bxzstr): a gzipped Themisto file works unchanged."""
    import gzip
    import shutil
    wl = synth.generate(800, 60, 5, n_present=2, n_templates=20, seed=4)
    paths = synth.write_themisto(str(tmp_path / "aln"), wl, paired=True)
    gpath = str(tmp_path / "g.txt")
    synth.write_grouping(gpath, wl)
    gz = []
    for p in paths:
        with open(p, "rb") as fi, gzip.open(p + ".gz", "wb") as fo:
            shutil.copyfileobj(fi, fo)
        gz.append(p + ".gz")
    R, T, rp, tg = dump(tmp_path, gz, gpath)
    assert np.array_equal(rp, wl.row_ptr) and np.array_equal(tg, wl.targets)


def test_rate_helper_matches_the_oracle(oracle, tmp_path):
    """b200::dirichlet_kld (host shim) takes theta and the aligned-read total; the oracle follows
    Sample::dirichlet_kld (src/Sample.cpp:99-131) literally, read by read, from the K x N posteriors."""
    rng = np.random.default_rng(5)
    K, N = 7, 300
    gamma = rng.normal(0, 3, size=(K, N))
    gamma -= np.log(np.exp(gamma).sum(axis=0))
    counts = rng.integers(1, 40, size=N).astype(np.float64)
    theta = (np.exp(gamma) * counts).sum(axis=1) / counts.sum()
    ref_log_kld, ref_rate = oracle.dirichlet_kld(gamma, np.log(counts))
    src = tmp_path / "rate.cpp"
    src.write_text('#include "msweep_b200.hpp"\n#include <cstdio>\n#include <cstdlib>\n'
                   'int main(int argc, char **argv) { std::vector<double> t; for (int i = 2; i < argc; ++i) t.push_back(std::atof(argv[i]));\n'
                   '  auto r = b200::dirichlet_kld(t, std::atof(argv[1]));\n'
                   '  for (size_t k = 0; k < t.size(); ++k) std::printf("%.17g %.17g\\n", r.first[k], r.second[k]); return 0; }\n')
    exe = tmp_path / "rate"
    lib = os.path.join(ROOT, "msweep_b200", "lib")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "msweep_b200", "host"),
                    str(src), "-o", str(exe), "-L", lib, "-lmsweep_b200", f"-Wl,-rpath,{lib}"], check=True)
    out = subprocess.run([str(exe), repr(float(counts.sum())), *[repr(float(x)) for x in theta]], capture_output=True, text=True, check=True)
    got = np.array([[float(x) for x in line.split()] for line in out.stdout.splitlines()])
    assert np.allclose(got[:, 0], ref_log_kld, rtol=1e-9, atol=1e-9)
    assert np.allclose(got[:, 1], ref_rate, rtol=1e-9)
