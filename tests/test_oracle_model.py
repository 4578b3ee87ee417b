"""CPU: model-level invariants of the oracle's optimiser.  The arithmetic of rcgpar is off-tree and
unpinned (oracle/oracle.hpp header), so these are properties ANY correct implementation of the model
must satisfy, plus agreement with the frozen fixture."""
import os

import numpy as np
import pytest

from msweep_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def case(oracle):
    wl = synth.generate(3000, 90, 6, n_present=3, n_templates=40, p_noise=0.03, seed=5)
    ec = oracle.ec_build_csr(wl.n_reads, wl.n_targets, wl.row_ptr, wl.targets)
    lik = oracle.lik_build(ec, wl.group_of_target, wl.group_sizes)
    return wl, ec, lik


def test_frozen_fixture_still_reproduced(oracle):
    g = np.load(os.path.join(GOLD, "small_case.npz"))
    ec = oracle.ec_build_csr(int(g["n_reads"]), int(g["n_targets"]), g["row_ptr"], g["targets"])
    assert np.array_equal(ec.hash, g["ec_hash"]) and np.array_equal(ec.count, g["ec_count"])
    assert np.array_equal(ec.rep_read, g["ec_rep_read"]) and np.array_equal(ec.pat_targets, g["ec_pat_targets"])
    lik = oracle.lik_build(ec, g["group_of_target"], g["group_sizes"], keep_hit_counts=True)
    assert np.array_equal(lik.hit_counts, g["hit_counts"])
    assert np.array_equal(lik.logl, g["logl"])
    r = oracle.vi_run("rcg", lik.logl, lik.log_counts)
    assert r.iters == int(g["rcg_iters"])
    assert np.max(np.abs(r.theta - g["rcg_theta"])) < 1e-12
    assert r.bound == pytest.approx(float(g["rcg_bound"]), rel=1e-12)
    assert np.array_equal(oracle.bootstrap_resample(ec.count, seed=int(g["boot_seed"]), n_replicates=3), g["boot_counts"])


def test_rcg_basic_properties(oracle, case):
    _, _, lik = case
    r = oracle.vi_run("rcg", lik.logl, lik.log_counts, want_gamma=True)
    assert r.converged and r.iters < 5000
    assert r.theta.sum() == pytest.approx(1.0, abs=1e-12)
    # theta = (N_k - alpha0) / sum c  (mixture_components identity)
    total = np.exp(lik.log_counts).sum()
    assert np.max(np.abs((r.N_k - 1.0) / total - r.theta)) < 1e-12
    # returned log-posteriors are normalised per class
    assert np.max(np.abs(np.exp(r.gamma).sum(axis=0) - 1.0)) < 1e-12
    # accepted iterations never decrease the bound (restarts exist for exactly that)
    acc = r.trace_bound
    assert np.all(np.diff(acc) > -1e-7 * np.abs(acc[:-1]).max())


def test_em_monotone_and_same_optimum_as_rcg(oracle, case):
    _, _, lik = case
    em = oracle.vi_run("em", lik.logl, lik.log_counts, tol=1e-10, max_iters=20000)
    rcg = oracle.vi_run("rcg", lik.logl, lik.log_counts, tol=1e-10, max_iters=20000)
    assert np.all(np.diff(em.trace_bound) >= -1e-9)
    assert abs(em.bound - rcg.bound) < 1e-4
    assert np.max(np.abs(em.theta - rcg.theta)) < 1e-4


def test_permutation_invariance(oracle, case):
    _, _, lik = case
    rng = np.random.default_rng(0)
    p = rng.permutation(lik.n_ecs)
    a = oracle.vi_run("em", lik.logl, lik.log_counts, tol=1e-9)
    b = oracle.vi_run("em", lik.logl[:, p], lik.log_counts[p], tol=1e-9)
    assert np.max(np.abs(a.theta - b.theta)) < 1e-9


def test_splitting_a_class_changes_nothing(oracle, case):
    _, _, lik = case
    j = int(np.argmax(lik.log_counts))
    c = np.exp(lik.log_counts[j])
    assert c >= 2
    logl = np.concatenate([lik.logl, lik.logl[:, j:j + 1]], axis=1)
    lc = np.concatenate([lik.log_counts, [np.log(1.0)]])
    lc[j] = np.log(c - 1.0)
    a = oracle.vi_run("em", lik.logl, lik.log_counts, tol=1e-9)
    b = oracle.vi_run("em", logl, lc, tol=1e-9)
    assert np.max(np.abs(a.theta - b.theta)) < 1e-9


def test_closed_forms(oracle):
    rng = np.random.default_rng(1)
    N = 50
    lc = np.log(rng.integers(1, 9, size=N).astype(np.float64))
    one = rng.normal(-3, 1, size=(1, N))
    for algo in ("rcg", "em"):
        assert oracle.vi_run(algo, one, lc).theta[0] == pytest.approx(1.0, abs=1e-14)
    # two identical groups: theta = 1/2, 1/2 under the symmetric prior.  (Only EM: with exactly symmetric
    # groups the Riemannian gradient norm is 0 and Fletcher-Reeves divides 0 by 0 — see DESIGN.md.)
    two = np.repeat(one, 2, axis=0)
    th = oracle.vi_run("em", two, lc).theta
    assert th[0] == pytest.approx(0.5, abs=1e-12) and th[1] == pytest.approx(0.5, abs=1e-12)


def test_rate_against_the_closed_form(oracle):
    """Sample::dirichlet_kld / get_rates (src/Sample.cpp:99-151) — in-tree formulas, checked against scipy."""
    from scipy.special import digamma, gammaln
    rng = np.random.default_rng(3)
    K, N = 5, 200
    gamma = rng.normal(0, 2, size=(K, N))
    gamma -= np.log(np.exp(gamma).sum(axis=0))
    counts = rng.integers(1, 30, size=N).astype(np.float64)
    log_kld, rate = oracle.dirichlet_kld(gamma, np.log(counts))
    a = (np.exp(gamma) * counts).sum(axis=1)
    a0 = a.sum()
    kld = np.maximum(gammaln(a0) - gammaln(a0 - a) - gammaln(a) + a * (digamma(a) - digamma(a0)), 1e-16)
    assert np.allclose(log_kld, np.log(kld), rtol=1e-8, atol=1e-8)
    top = max(0.0, np.log(kld).max())                      # the running maximum starts at 0 (:138-142)
    expect = np.exp(np.log(kld) - (np.log(np.exp(np.log(kld) - top).sum()) + top))
    assert np.allclose(rate, expect, rtol=1e-8)
    assert rate.sum() == pytest.approx(1.0, abs=1e-12)


def test_bin_rule(oracle):
    gamma = np.log(np.array([[0.9, 0.5, 0.05], [0.1, 0.5, 0.95]]))
    theta = np.array([0.5, 0.5])
    read_ptr, read_ids = np.array([0, 2, 3, 6], np.uint64), np.array([4, 9, 1, 0, 2, 7], np.uint32)
    bins = oracle.bin_reads(gamma, theta, np.ones(2, np.uint8), read_ptr, read_ids)
    assert bins[0].tolist() == [1, 4, 9] and bins[1].tolist() == [0, 1, 2, 7]
    assert oracle.bin_reads(gamma, theta, np.array([0, 1], np.uint8), read_ptr, read_ids)[0].size == 0


def test_zero_count_classes_are_ignored(oracle, case):
    """Bootstrap feeds log(0) = -inf for classes that were not resampled (src/BootstrapSample.cpp:67-72)."""
    _, _, lik = case
    lc = lik.log_counts.copy()
    drop = np.arange(lik.n_ecs) % 3 == 0
    lc[drop] = -np.inf
    a = oracle.vi_run("em", lik.logl, lc, tol=1e-9)
    b = oracle.vi_run("em", lik.logl[:, ~drop], lik.log_counts[~drop], tol=1e-9)
    assert np.all(np.isfinite(a.theta))
    assert np.max(np.abs(a.theta - b.theta)) < 1e-12
    assert a.bound == pytest.approx(b.bound, rel=1e-12)
    # RCG: the gradient norm is not count-weighted, so unobserved classes still steer the conjugate
    # direction; same optimum, slightly different path
    a = oracle.vi_run("rcg", lik.logl, lc)
    b = oracle.vi_run("rcg", lik.logl[:, ~drop], lik.log_counts[~drop])
    assert np.all(np.isfinite(a.theta)) and np.isfinite(a.bound)
    assert np.max(np.abs(a.theta - b.theta)) < 1e-5


def test_min_hits_keeps_the_estimates(oracle, case):
    """README.md:136-140: pruning groups without hits changes nothing beyond the prior mass."""
    wl, ec, lik = case
    lik1 = oracle.lik_build(ec, wl.group_of_target, wl.group_sizes, min_hits=1)
    assert lik1.n_groups <= lik.n_groups
    a = oracle.vi_run("rcg", lik.logl, lik.log_counts, tol=1e-9)
    b = oracle.vi_run("rcg", lik1.logl, lik1.log_counts, tol=1e-9)
    kept = lik1.mask.astype(bool)
    assert np.max(np.abs(a.theta[kept] - b.theta)) < 5e-3
    assert a.theta[~kept].sum() < 5e-3


def test_themisto_text_roundtrip(oracle, tmp_path):
    """include/mSWEEP_alignment.hpp:54-135: paired files, unordered lines, intersection merge."""
    wl = synth.generate(300, 40, 4, n_present=2, n_templates=10, seed=9)
    paths = synth.write_themisto(str(tmp_path / "aln"), wl, paired=True, shuffle_frac=0.1)
    n, rp, tg = oracle.read_themisto(paths, wl.n_targets, "intersection")
    assert n == wl.n_reads
    assert np.array_equal(rp, wl.row_ptr) and np.array_equal(tg, wl.targets)
    n, rp_u, tg_u = oracle.read_themisto(paths, wl.n_targets, "union")
    assert rp_u[-1] >= rp[-1]
    with pytest.raises(RuntimeError, match="Unrecognized option"):
        oracle.read_themisto(paths, wl.n_targets, "unpaired")
    bad = tmp_path / "bad.aln"
    bad.write_text("0 1 2\nx 3\n")
    with pytest.raises(RuntimeError, match="line 2"):
        oracle.read_themisto([str(bad)], 40, "intersection")
