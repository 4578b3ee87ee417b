"""GPU: the CUDA path, called through the C ABI, against the CPU oracle on the same seeded inputs and
against the frozen fixtures.  Integer artefacts are compared bit-exactly; floating point within the
tolerances BASELINE.json's north_star states (theta 1e-6 absolute, ELBO 1e-9 relative, fp64)."""
import os

import numpy as np
import pytest

from msweep_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
THETA_TOL = 1e-6      # north_star: abundances within 1e-6 absolute
ELBO_RTOL = 1e-9      # north_star: ELBO within 1e-9 relative (fp64)


def _ec_equal(dev, ref):
    assert np.array_equal(dev.hash, ref.hash)
    assert np.array_equal(dev.count, ref.count)
    assert np.array_equal(dev.rep_read, ref.rep_read)
    assert np.array_equal(dev.pat_ptr, ref.pat_ptr)
    assert np.array_equal(dev.pat_targets, ref.pat_targets)
    assert np.array_equal(dev.read_ptr, ref.read_ptr)
    assert np.array_equal(dev.read_ids, ref.read_ids)


CASES = {
    "tiny": dict(n_reads=64, n_targets=12, n_groups=3, n_present=2, n_templates=4, p_noise=0.1, seed=1),
    "small": dict(n_reads=5000, n_targets=150, n_groups=10, n_present=3, n_templates=80, p_noise=0.05, seed=2),
    "oddK": dict(n_reads=8000, n_targets=301, n_groups=7, n_present=4, n_templates=100, p_noise=0.02, seed=3),
    "wideK": dict(n_reads=6000, n_targets=3000, n_groups=300, n_present=6, n_templates=60, p_noise=0.02, seed=4),
    "c1_like": dict(n_reads=60000, n_targets=3000, n_groups=50, n_present=5, n_templates=400, p_noise=0.02, seed=5),
}


@pytest.fixture(scope="module", params=list(CASES))
def case(request, oracle, mswb, ctx):
    wl = synth.generate(**CASES[request.param])
    ec = oracle.ec_build_csr(wl.n_reads, wl.n_targets, wl.row_ptr, wl.targets)
    aln = mswb.Alignment(ctx, wl.n_reads, wl.n_targets, wl.row_ptr, wl.targets)
    return request.param, wl, ec, aln


def test_ec_table_bit_exact(case):
    _, wl, ec, aln = case
    assert aln.n_ecs == ec.n_ecs and aln.n_reads == wl.n_reads
    assert aln.n_aligned == int(ec.count.sum())
    _ec_equal(aln.export(), ec)


def test_hit_counts_and_logl(case, oracle, mswb, ctx):
    _, wl, ec, aln = case
    ref = oracle.lik_build(ec, wl.group_of_target, wl.group_sizes, keep_hit_counts=True)
    lik = mswb.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes)
    assert (lik.n_groups_all, lik.n_groups, lik.n_ecs) == (wl.n_groups, ref.n_groups, ec.n_ecs)
    assert np.array_equal(lik.export_hit_counts(), ref.hit_counts)          # integers: bit-exact
    got = lik.export_logl()
    assert np.array_equal(got, ref.logl)       # same libm calls on the host-built table: identical doubles
    assert np.all(lik.mask() == 1)


@pytest.mark.parametrize("min_hits", [1, 3, 50])
def test_min_hits_mask(case, oracle, mswb, ctx, min_hits):
    _, wl, ec, aln = case
    ref = oracle.lik_build(ec, wl.group_of_target, wl.group_sizes, min_hits=min_hits)
    if ref.n_groups == 0:
        with pytest.raises(mswb.MswbError, match="removed every group"):
            mswb.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes, min_hits=min_hits)
        return
    lik = mswb.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes, min_hits=min_hits)
    mask, hits = lik.mask(want_hits=True)
    assert np.array_equal(hits, ref.hits)
    assert np.array_equal(mask, ref.mask)
    assert lik.n_groups == ref.n_groups
    assert np.array_equal(lik.export_logl(), ref.logl)


@pytest.mark.parametrize("algo", ["rcg", "em"])
def test_vi_parity(case, oracle, mswb, ctx, algo):
    name, wl, ec, aln = case
    ref_l = oracle.lik_build(ec, wl.group_of_target, wl.group_sizes)
    lik = mswb.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes)
    ref = oracle.vi_run(algo, ref_l.logl, ref_l.log_counts, tol=1e-6, max_iters=5000)
    got = lik.vi_run(mswb.ALGO_RCG if algo == "rcg" else mswb.ALGO_EM, tol=1e-6, max_iters=5000)
    assert got.converged == ref.converged
    assert got.iters == ref.iters, (got.iters, ref.iters)
    assert np.max(np.abs(got.theta - ref.theta)) < THETA_TOL
    assert abs(got.bound - ref.bound) <= ELBO_RTOL * abs(ref.bound)
    assert got.theta.sum() == pytest.approx(1.0, abs=1e-12)


def test_rcg_trajectory_and_posteriors(case, oracle, mswb, ctx):
    """Same trajectory, not just same optimum: bound per iteration, restarts, final log-posteriors."""
    name, wl, ec, aln = case
    ref_l = oracle.lik_build(ec, wl.group_of_target, wl.group_sizes)
    lik = mswb.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes)
    ref = oracle.vi_run("rcg", ref_l.logl, ref_l.log_counts, tol=1e-6, max_iters=40, want_gamma=True)
    s = lik.vi_begin(mswb.ALGO_RCG, tol=1e-6, max_iters=40)
    s.step(40)
    tb, tg, tr = s.trace()
    got = s.finish()
    assert got.iters == ref.iters
    assert np.array_equal(tr, ref.trace_reset)
    assert np.max(np.abs(tb - ref.trace_bound) / np.abs(ref.trace_bound)) < ELBO_RTOL
    assert np.allclose(tg, ref.trace_gnorm, rtol=1e-6, atol=1e-9)
    gam = lik.posteriors()
    big = ref.gamma > -30            # log-posteriors of any weight; far tails are exp-underflow noise
    assert np.max(np.abs(gam[big] - ref.gamma[big])) < 1e-6      # i.e. responsibilities agree to 1e-6 RELATIVE
    assert np.max(np.abs(np.exp(gam) - np.exp(ref.gamma))) < 1e-6      # 40 CG steps amplify rounding; theta's tolerance
    assert np.max(np.abs(got.N_k - ref.N_k)) < 1e-6 * max(1.0, ref.N_k.max())


def test_em_posteriors(case, oracle, mswb, ctx):
    name, wl, ec, aln = case
    ref_l = oracle.lik_build(ec, wl.group_of_target, wl.group_sizes)
    lik = mswb.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes)
    ref = oracle.vi_run("em", ref_l.logl, ref_l.log_counts, tol=1e-6, max_iters=25, want_gamma=True)
    got = lik.vi_run(mswb.ALGO_EM, tol=1e-6, max_iters=25)
    assert got.iters == ref.iters
    gam = lik.posteriors()
    assert np.max(np.abs(np.exp(gam) - np.exp(ref.gamma))) < 1e-9


@pytest.mark.parametrize("algo", ["rcg", "em"])
def test_read_bins_follow_the_threshold_rule(case, oracle, mswb, ctx, algo):
    """mswb_vi_assign (the --bin-reads hand-off): a class, with all of its reads, joins group k's bin when its
    log-posterior reaches log(theta_k).  Same posteriors in, identical bins out (integers: bit-exact)."""
    name, wl, ec, aln = case
    lik = mswb.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes)
    got = lik.vi_run(mswb.ALGO_RCG if algo == "rcg" else mswb.ALGO_EM, tol=1e-6, max_iters=30)
    gam = lik.posteriors()
    K = lik.n_groups
    want = np.ones(K, np.uint8)
    want[1::3] = 0
    with np.errstate(divide="ignore"):
        thr = np.where(want == 1, np.log(got.theta), np.inf)
    bins = lik.assign(aln, thr)
    ref = oracle.bin_reads(gam, got.theta, want, ec.read_ptr, ec.read_ids)
    assert len(bins) == K
    for k in range(K):
        assert np.array_equal(bins[k], ref[k]), (name, k)
        if not want[k]:
            assert bins[k].size == 0
    assert sum(b.size for b in bins) > 0
    # no threshold at all: every group takes every aligned read, ascending
    everything = lik.assign(aln, np.full(K, -np.inf))
    all_reads = np.sort(ec.read_ids)
    assert all(np.array_equal(b, all_reads) for b in everything)
    # the oracle's own run of the optimiser lands on the same bins up to classes sitting on a threshold
    ref_l = oracle.lik_build(ec, wl.group_of_target, wl.group_sizes)
    r = oracle.vi_run(algo, ref_l.logl, ref_l.log_counts, tol=1e-6, max_iters=30, want_gamma=True)
    own = oracle.bin_reads(r.gamma, r.theta, want, ec.read_ptr, ec.read_ids)
    moved = sum(np.setxor1d(bins[k], own[k]).size for k in range(K))
    assert moved <= 1e-3 * max(1, sum(b.size for b in own))


def test_fp32_storage_tolerance(case, oracle, mswb, ctx):
    """fp32 storage of the linear-domain likelihood (EM only), fp64 accumulation across classes.
    Tolerance stated separately, as north_star asks.  Compared at a FIXED iteration count: the ELBO
    carries ~1e-7 relative storage noise, so a 1e-6 absolute stopping rule fires at a different
    iteration than in fp64 and EM moves theta by ~1e-5 per late iteration.
    fp32-storage tolerance: theta 2e-6 absolute, ELBO 1e-6 relative."""
    name, wl, ec, aln = case
    ref_l = oracle.lik_build(ec, wl.group_of_target, wl.group_sizes)
    ref = oracle.vi_run("em", ref_l.logl, ref_l.log_counts, tol=0.0, max_iters=40)
    lik = mswb.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes, storage=mswb.STORE_F32)
    got = lik.vi_run(mswb.ALGO_EM, tol=0.0, max_iters=40)
    assert got.iters == ref.iters == 40
    assert np.max(np.abs(got.theta - ref.theta)) < 2e-6
    assert abs(got.bound - ref.bound) <= 1e-6 * abs(ref.bound)
    with pytest.raises(mswb.MswbError, match="RCG needs"):
        lik.vi_run(mswb.ALGO_RCG)


@pytest.mark.parametrize("min_hits", [0, 3])
def test_sparse_storage_is_lossless(case, oracle, mswb, ctx, min_hits):
    """MSWB_STORE_SPARSE keeps, per class, log(zero_inflation) once plus its (group, value) hits in fp64: the same
    numbers as the dense fp64 matrix, so EM must follow the oracle's fp64 trajectory (fp64 tolerances apply)."""
    name, wl, ec, aln = case
    ref_l = oracle.lik_build(ec, wl.group_of_target, wl.group_sizes, min_hits=min_hits)
    if ref_l.n_groups == 0:
        pytest.skip("all groups pruned")
    ref = oracle.vi_run("em", ref_l.logl, ref_l.log_counts, tol=1e-6, max_iters=5000)
    lik = mswb.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes, min_hits=min_hits, storage=mswb.STORE_SPARSE)
    assert lik.n_groups == ref_l.n_groups
    got = lik.vi_run(mswb.ALGO_EM, tol=1e-6, max_iters=5000)
    assert got.iters == ref.iters and got.converged == ref.converged
    assert np.max(np.abs(got.theta - ref.theta)) < THETA_TOL
    assert abs(got.bound - ref.bound) <= ELBO_RTOL * abs(ref.bound)
    dense = mswb.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes, min_hits=min_hits).vi_run(mswb.ALGO_EM, tol=0.0, max_iters=25)
    sparse = lik.vi_run(mswb.ALGO_EM, tol=0.0, max_iters=25)
    assert np.max(np.abs(dense.theta - sparse.theta)) < 1e-12
    # bootstrap replicates run on the sparse form too
    thetas, _ = lik.bootstrap_run(2, seed=5, algo=mswb.ALGO_EM, max_iters=30, tol=0.0)
    counts = oracle.bootstrap_resample(ec.count, 5, 2)
    for r in range(2):
        with np.errstate(divide="ignore"):
            want = oracle.vi_run("em", ref_l.logl, np.log(counts[r].astype(np.float64)), tol=0.0, max_iters=30)
        assert np.max(np.abs(thetas[r] - want.theta)) < THETA_TOL


@pytest.mark.parametrize("min_hits", [0, 3])
def test_sparse_rcg_follows_the_dense_trajectory(case, oracle, mswb, ctx, min_hits):
    """RCG on the sparse storage (vi_sparse_rcg.cuh): off a class's hits gamma and the search direction are exactly
    a_k + b_j, so the optimiser runs on two K-vectors, two N-vectors and the hits — the same iteration as the dense sweeps,
    step for step: the oracle's iteration count, restart pattern, bound trajectory and abundances (fp64 tolerances)."""
    name, wl, ec, aln = case
    ref_l = oracle.lik_build(ec, wl.group_of_target, wl.group_sizes, min_hits=min_hits)
    if ref_l.n_groups < 2:
        pytest.skip("needs two groups")
    ref = oracle.vi_run("rcg", ref_l.logl, ref_l.log_counts, tol=1e-6, max_iters=5000)
    lik = mswb.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes, min_hits=min_hits, storage=mswb.STORE_SPARSE)
    sess = lik.vi_begin(mswb.ALGO_RCG, tol=1e-6, max_iters=5000)
    while True:
        sess.step(8)
        st = sess.poll()
        if st.converged or st.iters >= 5000:
            break
    tb, tg, tr = sess.trace()
    got = sess.finish()
    assert got.iters == ref.iters and got.converged == ref.converged
    assert np.array_equal(tr, ref.trace_reset)
    assert np.max(np.abs(tb - ref.trace_bound) / np.abs(ref.trace_bound)) <= ELBO_RTOL
    assert np.allclose(tg, ref.trace_gnorm, rtol=1e-6, atol=1e-9 * (1.0 + ref.trace_gnorm[0]))
    assert np.max(np.abs(got.theta - ref.theta)) < THETA_TOL
    assert abs(got.theta.sum() - 1.0) < 1e-12
    # bit-reproducible (fixed-point scatter, fixed-order sums), and a bootstrap replicate with unobserved classes
    again = lik.vi_run(mswb.ALGO_RCG, tol=1e-6, max_iters=5000)
    assert np.array_equal(again.theta, got.theta) and again.bound == got.bound
    thetas, iters = lik.bootstrap_run(2, seed=5)
    counts = oracle.bootstrap_resample(ec.count, 5, 2)
    for r in range(2):
        with np.errstate(divide="ignore"):
            want = oracle.vi_run("rcg", ref_l.logl, np.log(counts[r].astype(np.float64)))
        assert iters[r] == want.iters
        assert np.max(np.abs(thetas[r] - want.theta)) < THETA_TOL


def test_fused_small_problem_kernel_equals_the_launch_per_sweep_path(case, mswb, ctx, monkeypatch):
    """Small problems on one GPU run whole RCG iterations inside one cooperative launch (rcgs_fused_kernel: the sweeps of
    vi_sparse_rcg.cuh between grid rendezvous).  Same code, same summation orders: the result must equal the three-launch
    path bit for bit — trajectory, restarts, abundances — however the iterations are cut into launches."""
    name, wl, ec, aln = case
    if len(wl.group_sizes) < 2:
        pytest.skip("needs two groups")
    lik = mswb.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes, storage=mswb.STORE_SPARSE)
    runs = {}
    for label, fused, chunk in (("launches", "0", 8), ("fused", "1", 8), ("fused_1", "1", 1), ("fused_100", "1", 100)):
        monkeypatch.setenv("MSWB_FUSED", fused)
        n0 = mswb.launch_count()
        sess = lik.vi_begin(mswb.ALGO_RCG, tol=1e-6, max_iters=300)
        while True:
            sess.step(chunk)
            st = sess.poll()
            if st.converged or st.iters >= 300:
                break
        tb, tg, tr = sess.trace()
        r = sess.finish()
        runs[label] = (tb, tg, tr, r, mswb.launch_count() - n0)
    tb0, tg0, tr0, r0, n_launches = runs["launches"]
    for label in ("fused", "fused_1", "fused_100"):
        tb, tg, tr, r, n = runs[label]
        assert r.iters == r0.iters and r.converged == r0.converged and r.resets == r0.resets, label
        assert np.array_equal(tb, tb0) and np.array_equal(tg, tg0) and np.array_equal(tr, tr0), label
        assert np.array_equal(r.theta, r0.theta) and r.bound == r0.bound, label
    assert runs["fused_100"][4] < n_launches / 3            # the iterations really ran inside few launches
    monkeypatch.setenv("MSWB_FUSED", "1")
    # the restart sweep inside the fused loop: a start that overshoots (large prior counts make early steps lose ground)
    a0 = np.full(lik.n_groups, 50.0)
    monkeypatch.setenv("MSWB_FUSED", "0")
    want = lik.vi_run(mswb.ALGO_RCG, alpha0=a0, tol=1e-6, max_iters=300)
    monkeypatch.setenv("MSWB_FUSED", "1")
    got = lik.vi_run(mswb.ALGO_RCG, alpha0=a0, tol=1e-6, max_iters=300)
    assert got.iters == want.iters and got.resets == want.resets and np.array_equal(got.theta, want.theta)


@pytest.mark.parametrize("algo", ["rcg", "em"])
def test_sparse_posteriors_and_bins_equal_the_dense_ones(case, oracle, mswb, ctx, algo):
    """Posterior export (Sample::store_probs, src/Sample.cpp:63-85) and the binning hand-off from the sparse storage:
    gamma = a_k + b_j off the hits (RCG) / l0 + digamma(N_k) - L_j (EM), explicit on the hits — the dense matrix, tile by tile."""
    name, wl, ec, aln = case
    code = mswb.ALGO_RCG if algo == "rcg" else mswb.ALGO_EM
    if algo == "rcg" and len(wl.group_sizes) < 2:
        pytest.skip("needs two groups")
    dense = mswb.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes)
    sparse = mswb.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes, storage=mswb.STORE_SPARSE)
    rd, rs = dense.vi_run(code), sparse.vi_run(code)
    assert rd.iters == rs.iters
    gd, gs = dense.posteriors(), sparse.posteriors()
    assert gd.shape == gs.shape
    assert np.max(np.abs(gd - gs)) < 1e-8
    assert np.max(np.abs(np.exp(gs).sum(axis=0) - 1.0)) < 1e-12
    n = min(7, sparse.n_ecs)
    assert np.array_equal(sparse.posteriors(3 if sparse.n_ecs > 10 else 0, (3 if sparse.n_ecs > 10 else 0) + n),
                          gs[:, (3 if sparse.n_ecs > 10 else 0):(3 if sparse.n_ecs > 10 else 0) + n])
    with np.errstate(divide="ignore"):
        thr = np.log(rs.theta) + 1e-6           # (away from ties: the two forms differ in the last bits)
    bd, bs = dense.assign(aln, thr), sparse.assign(aln, thr)
    assert all(np.array_equal(x, y) for x, y in zip(bd, bs))


def test_golden_fixture(mswb, ctx):
    """The frozen numbers (no oracle at run time)."""
    g = np.load(os.path.join(GOLD, "small_case.npz"))
    aln = mswb.Alignment(ctx, int(g["n_reads"]), int(g["n_targets"]), g["row_ptr"], g["targets"])
    e = aln.export()
    assert np.array_equal(e.hash, g["ec_hash"]) and np.array_equal(e.count, g["ec_count"])
    assert np.array_equal(e.rep_read, g["ec_rep_read"]) and np.array_equal(e.pat_targets, g["ec_pat_targets"])
    assert np.array_equal(e.read_ids, g["ec_read_ids"])
    lik = mswb.Likelihood.build(ctx, aln, g["group_of_target"], g["group_sizes"])
    assert np.array_equal(lik.export_hit_counts(), g["hit_counts"])
    assert np.array_equal(lik.export_logl(), g["logl"])
    lik1 = mswb.Likelihood.build(ctx, aln, g["group_of_target"], g["group_sizes"], min_hits=1)
    m, h = lik1.mask(want_hits=True)
    assert np.array_equal(m, g["mask_minhits1"]) and np.array_equal(h, g["hits_minhits1"])
    r = lik.vi_run(mswb.ALGO_RCG)
    assert r.iters == int(g["rcg_iters"])
    assert np.max(np.abs(r.theta - g["rcg_theta"])) < THETA_TOL
    assert abs(r.bound - float(g["rcg_bound"])) <= ELBO_RTOL * abs(float(g["rcg_bound"]))
    em = lik.vi_run(mswb.ALGO_EM)
    assert em.iters == int(g["em_iters"])
    assert np.max(np.abs(em.theta - g["em_theta"])) < THETA_TOL
    assert np.array_equal(lik.bootstrap_resample(int(g["boot_seed"]), 3), g["boot_counts"])


# ---- edge cases ---------------------------------------------------------------------------------
def test_empty_and_degenerate_inputs(oracle, mswb, ctx):
    # all reads unaligned
    aln = mswb.Alignment(ctx, 5, 9, np.zeros(6, np.uint64), np.zeros(0, np.uint32))
    assert (aln.n_ecs, aln.n_aligned, aln.n_reads) == (0, 0, 5)
    with pytest.raises(mswb.MswbError, match="no read aligned"):
        mswb.Likelihood.build(ctx, aln, np.zeros(9, np.uint32), np.array([9], np.uint64))
    # zero reads
    aln = mswb.Alignment(ctx, 0, 9, np.zeros(1, np.uint64), np.zeros(0, np.uint32))
    assert aln.n_ecs == 0
    # a single read, single target
    aln = mswb.Alignment(ctx, 1, 9, np.array([0, 1], np.uint64), np.array([4], np.uint32))
    e = aln.export()
    assert list(e.hash) == [oracle.pattern_hash([4])] and list(e.count) == [1]
    # unsorted row is rejected, not silently mis-hashed
    with pytest.raises(mswb.MswbError, match="ascending"):
        mswb.Alignment(ctx, 1, 9, np.array([0, 2], np.uint64), np.array([5, 3], np.uint32))
    with pytest.raises(mswb.MswbError, match="ascending"):
        mswb.Alignment(ctx, 1, 9, np.array([0, 1], np.uint64), np.array([9], np.uint32))


def test_equal_patterns_far_apart_merge(oracle, mswb, ctx):
    rng = np.random.default_rng(3)
    rows = [sorted(rng.choice(500, size=int(n), replace=False).tolist()) for n in rng.integers(0, 30, size=4000)]
    for i in range(0, 4000, 7):
        rows[i] = rows[(i * 13) % 4000]
    ptr = np.cumsum([0] + [len(r) for r in rows]).astype(np.uint64)
    tg = np.array([t for r in rows for t in r], np.uint32)
    _ec_equal(mswb.Alignment(ctx, len(rows), 500, ptr, tg).export(), oracle.ec_build_csr(len(rows), 500, ptr, tg))


@pytest.mark.parametrize("K,N", [(1, 37), (2, 1), (3, 1000), (4, 3000), (5, 777), (8, 5000), (9, 300), (13, 2500), (16, 4097), (17, 999), (25, 1300), (31, 640), (32, 2048),
                                 (50, 3001), (64, 1025),
                                 (33, 513), (65, 2049), (100, 333), (129, 700), (190, 450), (257, 300),
                                 (330, 257), (400, 129), (460, 131), (560, 90), (600, 200), (900, 110), (1100, 150), (1300, 100), (1500, 60), (1600, 70), (1900, 65),
                                 (2100, 64), (2600, 40), (3100, 33), (3700, 20), (5000, 12), (9000, 9)])
def test_dense_entry_all_tile_shapes(oracle, mswb, ctx, K, N):
    """mswb_lik_from_dense over every compiled tile shape, ragged N, non-uniform prior, zero-count classes."""
    rng = np.random.default_rng(K * 1000 + N)
    logl = rng.normal(-6.0, 2.5, size=(K, N))
    logl[rng.integers(0, K, size=N), np.arange(N)] = rng.normal(-0.5, 0.2, size=N)   # every class has a likely group
    counts = rng.integers(1, 50, size=N).astype(np.float64)
    lc = np.log(counts)
    if N > 10:
        lc[::5] = -np.inf                       # bootstrap-style unobserved classes
    alpha0 = rng.uniform(0.5, 2.0, size=K)
    lik = mswb.Likelihood.from_dense(ctx, logl, np.where(np.isfinite(lc), lc, 0.0))
    assert np.array_equal(lik.export_logl(), logl)
    for algo, code in (("rcg", mswb.ALGO_RCG), ("em", mswb.ALGO_EM)):
        if algo == "rcg" and K == 1:
            continue
        ref = oracle.vi_run(algo, logl, lc, alpha0=alpha0, tol=1e-7, max_iters=30)
        got = lik.vi_run(code, alpha0=alpha0, log_counts=lc, tol=1e-7, max_iters=30)
        assert got.iters == ref.iters
        assert np.max(np.abs(got.theta - ref.theta)) < THETA_TOL
        assert abs(got.bound - ref.bound) <= ELBO_RTOL * abs(ref.bound)


@pytest.mark.parametrize("algo,K,env,value", [("em", 1000, "MSWB_EM_TMA", "1"), ("rcg", 1100, "MSWB_RCG_TMA", "0"),
                                              ("rcg", 1500, "MSWB_RCG_TMA", "0"), ("rcg", 2000, "MSWB_RCG_TMA", "0")])
def test_tma_stage_ring_variant(oracle, mswb, ctx, algo, K, env, value):
    """Two ways of feeding the SM are shipped: direct streaming loads and the cp.async.bulk + mbarrier
    stage ring.  Full-width RCG sweeps default to the ring, the EM sweep to direct loads; the environment switch selects
    the other one, which has to give the same answers as the default path and the oracle."""
    rng = np.random.default_rng(K)
    N = 700
    logl = rng.normal(-6.0, 2.0, size=(K, N))
    logl[rng.integers(0, K, size=N), np.arange(N)] = -0.3
    lc = np.log(rng.integers(1, 30, size=N).astype(np.float64))
    ref = oracle.vi_run(algo, logl, lc, tol=1e-7, max_iters=12)
    lik = mswb.Likelihood.from_dense(ctx, logl, lc)
    code = mswb.ALGO_RCG if algo == "rcg" else mswb.ALGO_EM
    default = lik.vi_run(code, tol=1e-7, max_iters=12)
    os.environ[env] = value
    try:
        other = lik.vi_run(code, tol=1e-7, max_iters=12)
    finally:
        del os.environ[env]
    assert other.iters == ref.iters == default.iters
    for got in (default, other):
        assert np.max(np.abs(got.theta - ref.theta)) < THETA_TOL and abs(got.bound - ref.bound) <= ELBO_RTOL * abs(ref.bound)
    assert np.max(np.abs(other.theta - default.theta)) < 1e-13


@pytest.mark.parametrize("K,N", [(3, 900), (7, 4000), (12, 1500), (30, 2600), (50, 3100), (100, 700), (128, 515)])
def test_short_rows_fp32_storage(oracle, mswb, ctx, K, N):
    """The sub-warp tile shapes of the fp32-stored EM sweep (rows of 1..32 float4 pieces) against the fp64 oracle at a fixed
    iteration count; fp32 STORAGE tolerance: theta 2e-6 absolute, ELBO 1e-6 relative (stated separately from fp64)."""
    rng = np.random.default_rng(K * 7 + N)
    logl = rng.normal(-5.0, 2.0, size=(K, N))
    logl[rng.integers(0, K, size=N), np.arange(N)] = rng.normal(-0.5, 0.2, size=N)
    lc = np.log(rng.integers(1, 50, size=N).astype(np.float64))
    ref = oracle.vi_run("em", logl, lc, tol=0.0, max_iters=20)
    got = mswb.Likelihood.from_dense(ctx, logl, lc, storage=mswb.STORE_F32).vi_run(mswb.ALGO_EM, tol=0.0, max_iters=20)
    assert got.iters == ref.iters == 20
    assert np.max(np.abs(got.theta - ref.theta)) < 2e-6
    assert abs(got.bound - ref.bound) <= 1e-6 * abs(ref.bound)


@pytest.mark.parametrize("algo", ["rcg", "em"])
@pytest.mark.parametrize("K,N", [(50, 30000), (100, 20000)])
def test_reduction_tail_variants(oracle, mswb, ctx, algo, K, N):
    """The per-CTA partial vectors are summed either by the last CTA of the sweep (which then also takes the control
    step: one launch per EM iteration) or by finalize_ctl_kernel; MSWB_TAIL_MAX=0 forces the latter.  Both must follow
    the oracle, and each other to summation-order noise."""
    rng = np.random.default_rng(K + N)
    logl = rng.normal(-6.0, 2.0, size=(K, N))
    logl[rng.integers(0, K, size=N), np.arange(N)] = -0.3
    lc = np.log(rng.integers(1, 30, size=N).astype(np.float64))
    ref = oracle.vi_run(algo, logl, lc, tol=1e-7, max_iters=15)
    lik = mswb.Likelihood.from_dense(ctx, logl, lc)
    code = mswb.ALGO_RCG if algo == "rcg" else mswb.ALGO_EM
    fused = lik.vi_run(code, tol=1e-7, max_iters=15)

    def launches_per_iteration():
        sess = lik.vi_begin(code, tol=-1e300 if algo == "rcg" else 0.0, max_iters=1000)
        sess.step(2)
        l0 = mswb.launch_count()
        sess.step(10)
        n = mswb.launch_count() - l0
        sess.finish()
        return n / 10.0

    per_iter_fused = launches_per_iteration()
    os.environ["MSWB_TAIL_MAX"] = "0"
    try:
        split = lik.vi_run(code, tol=1e-7, max_iters=15)
        per_iter_split = launches_per_iteration()
    finally:
        del os.environ["MSWB_TAIL_MAX"]
    assert fused.iters == split.iters == ref.iters
    for got in (fused, split):
        assert np.max(np.abs(got.theta - ref.theta)) < THETA_TOL and abs(got.bound - ref.bound) <= ELBO_RTOL * abs(ref.bound)
    assert np.max(np.abs(fused.theta - split.theta)) < 1e-10     # two summation orders, 15 iterations apart
    # launch budget per iteration on one GPU: EM 1 (fused) / 2, RCG 3 / 4 (sweep A, sweep B [, reduction], restart sweep)
    assert per_iter_fused == (1.0 if algo == "em" else 3.0), per_iter_fused
    assert per_iter_split == (2.0 if algo == "em" else 4.0), per_iter_split


def test_sparse_pass_is_bit_reproducible(mswb, ctx):
    """The sparse EM pass scatters into fixed-point accumulators (integer adds commute): two runs give identical bits."""
    wl = synth.generate_ec_patterns(200_000, 300, 12, n_present=20, seed=9)
    aln = mswb.Alignment(ctx, wl.n_reads, wl.n_targets, wl.row_ptr, wl.targets)
    lik = mswb.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes, storage=mswb.STORE_SPARSE)
    runs = [lik.vi_run(mswb.ALGO_EM, tol=0.0, max_iters=40) for _ in range(3)]
    for r in runs[1:]:
        assert np.array_equal(r.theta, runs[0].theta) and r.bound == runs[0].bound
    dense = mswb.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes).vi_run(mswb.ALGO_EM, tol=0.0, max_iters=40)
    assert np.max(np.abs(dense.theta - runs[0].theta)) < 1e-11
    assert abs(dense.bound - runs[0].bound) <= 1e-12 * abs(dense.bound)


def test_group_sizes_must_match_the_indicators(mswb, ctx):
    """The hit count of a class indexes the group's row of the lookup table: sizes that disagree with the indicators
    (a C-ABI caller's mistake; the reference's Grouping derives both from one file) are refused."""
    wl = synth.generate(2000, 60, 6, n_present=2, n_templates=20, seed=4)
    aln = mswb.Alignment(ctx, wl.n_reads, wl.n_targets, wl.row_ptr, wl.targets)
    bad = wl.group_sizes.copy()
    bad[2] -= 1
    with pytest.raises(mswb.MswbError, match="group_sizes"):
        mswb.Likelihood.build(ctx, aln, wl.group_of_target, bad)
    bad[2] = 0
    with pytest.raises(mswb.MswbError, match="group_sizes"):
        mswb.Likelihood.build(ctx, aln, wl.group_of_target, bad)


def test_bootstrap_counts_bit_exact(oracle, mswb, ctx):
    wl = synth.generate(20000, 200, 10, n_present=3, n_templates=300, seed=8)
    ec = oracle.ec_build_csr(wl.n_reads, wl.n_targets, wl.row_ptr, wl.targets)
    aln = mswb.Alignment(ctx, wl.n_reads, wl.n_targets, wl.row_ptr, wl.targets)
    lik = mswb.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes)
    for seed in (1, 42, -7):
        assert np.array_equal(lik.bootstrap_resample(seed, 4), oracle.bootstrap_resample(ec.count, seed, 4))
    assert np.array_equal(lik.bootstrap_resample(5, 2, bootstrap_count=777), oracle.bootstrap_resample(ec.count, 5, 2, 777))
    ph = lik.bootstrap_resample(5, 2, rng_mode=mswb.RNG_PHILOX)
    assert list(ph.sum(axis=1)) == [int(ec.count.sum())] * 2 and not np.array_equal(ph[0], ph[1])


@pytest.mark.parametrize("segments", ["2", "5", "16"])
def test_mt64_segments_equal_the_sequential_stream(oracle, mswb, ctx, monkeypatch, segments):
    """A replicate's std::mt19937_64 stream generated in parallel segments (bootstrap.cu: mt64_chain_kernel applies
    z^L mod phi to the generator's state, mt64_segments_kernel generates all segments at once) must be the sequential
    stream bit for bit: resampled counts equal libstdc++'s (the oracle calls std::mt19937_64 itself) over several
    replicates — the consumed count inside the generator's array moves from replicate to replicate — and for draw counts
    that are not multiples of 312 or of the segment length."""
    wl = synth.generate(20000, 200, 10, n_present=3, n_templates=300, seed=8)
    ec = oracle.ec_build_csr(wl.n_reads, wl.n_targets, wl.row_ptr, wl.targets)
    aln = mswb.Alignment(ctx, wl.n_reads, wl.n_targets, wl.row_ptr, wl.targets)
    lik = mswb.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes)
    monkeypatch.setenv("MSWB_MT_SEGMENTS", segments)
    assert np.array_equal(lik.bootstrap_resample(3, 5), oracle.bootstrap_resample(ec.count, 3, 5))
    for count in (313, 1000, 6241, 50001):
        assert np.array_equal(lik.bootstrap_resample(-9, 3, bootstrap_count=count), oracle.bootstrap_resample(ec.count, -9, 3, count)), count
    # the replicate loop on top of it, with ranks jumping over each other's replicates
    lik_sp = mswb.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes, storage=mswb.STORE_SPARSE)
    monkeypatch.setenv("MSWB_MT_SEGMENTS", "0")
    whole, it_whole = lik_sp.bootstrap_run(6, seed=21)
    monkeypatch.setenv("MSWB_MT_SEGMENTS", segments)
    again, it_again = lik_sp.bootstrap_run(6, seed=21)
    assert np.array_equal(whole, again) and list(it_whole) == list(it_again)
    monkeypatch.setenv("MSWB_MT_JUMP", "1")
    t, _ = lik_sp.bootstrap_run(6, seed=21, replica_rank=1, replica_world=3)
    assert np.array_equal(t[[1, 4]], whole[[1, 4]])


def test_philox_bootstrap_is_a_fair_multinomial(mswb, ctx):
    """MSWB_RNG_PHILOX has no reference stream to match: check it statistically (counts within 6 sigma of n p, replicates
    and seeds independent, totals exact)."""
    wl = synth.generate(30000, 200, 10, n_present=3, n_templates=40, seed=18)
    aln = mswb.Alignment(ctx, wl.n_reads, wl.n_targets, wl.row_ptr, wl.targets)
    lik = mswb.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes)
    c = aln.export().count.astype(np.float64)
    n, p = c.sum(), c / c.sum()
    reps = lik.bootstrap_resample(123, 6, rng_mode=mswb.RNG_PHILOX).astype(np.float64)
    assert np.all(reps.sum(axis=1) == n)
    sigma = np.sqrt(n * p * (1 - p)) + 1e-9
    assert np.max(np.abs(reps - n * p) / np.maximum(sigma, 1.0)) < 6.0
    assert np.max(np.abs(reps.mean(axis=0) - n * p) / np.maximum(sigma / np.sqrt(6), 1.0)) < 6.0
    other = lik.bootstrap_resample(124, 1, rng_mode=mswb.RNG_PHILOX)[0]
    assert not np.array_equal(other, reps[0]) and not np.array_equal(reps[0], reps[1])
    again = lik.bootstrap_resample(123, 6, rng_mode=mswb.RNG_PHILOX)
    assert np.array_equal(again, reps.astype(np.uint32))               # counter-based: reproducible for a seed


def test_bootstrap_run_matches_reference_loop(oracle, mswb, ctx):
    """src/mSWEEP.cpp:496-518: resample, re-estimate from a cold start, once per replicate."""
    wl = synth.generate(8000, 120, 6, n_present=3, n_templates=80, seed=12)
    ec = oracle.ec_build_csr(wl.n_reads, wl.n_targets, wl.row_ptr, wl.targets)
    ref_l = oracle.lik_build(ec, wl.group_of_target, wl.group_sizes)
    aln = mswb.Alignment(ctx, wl.n_reads, wl.n_targets, wl.row_ptr, wl.targets)
    lik = mswb.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes)
    B = 5
    thetas, iters = lik.bootstrap_run(B, seed=99)
    counts = oracle.bootstrap_resample(ec.count, 99, B)
    for r in range(B):
        with np.errstate(divide="ignore"):
            ref = oracle.vi_run("rcg", ref_l.logl, np.log(counts[r].astype(np.float64)))
        assert iters[r] == ref.iters
        assert np.max(np.abs(thetas[r] - ref.theta)) < THETA_TOL
    # replicas spread over two "GPUs": each computes its own rows, the union equals the single run
    t0, _ = lik.bootstrap_run(B, seed=99, replica_rank=0, replica_world=2)
    t1, _ = lik.bootstrap_run(B, seed=99, replica_rank=1, replica_world=2)
    merged = np.where(np.isnan(t0), t1, t0)
    assert np.array_equal(merged, thetas)


@pytest.mark.parametrize("algo_name,storage", [("rcg", "sparse"), ("em", "f64")])
def test_bootstrap_ranks_jump_over_each_others_draws(mswb, ctx, monkeypatch, algo_name, storage):
    """One std::mt19937_64 serves all replicates (src/BootstrapSample.cpp:48-60).  A rank that owns every third replicate
    either produces and drops the draws in between (MSWB_MT_JUMP=0) or jumps the generator over them (mt64_jump.cu):
    the replicates must come out bit-identical, and equal to the single-rank run."""
    wl = synth.generate(9000, 90, 7, n_present=3, n_templates=60, seed=31)
    aln = mswb.Alignment(ctx, wl.n_reads, wl.n_targets, wl.row_ptr, wl.targets)
    lik = mswb.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes,
                                storage=mswb.STORE_SPARSE if storage == "sparse" else mswb.STORE_F64)
    algo = mswb.ALGO_RCG if algo_name == "rcg" else mswb.ALGO_EM
    B, W = 8, 3
    whole, it_whole = lik.bootstrap_run(B, seed=-5, algo=algo)
    for mode in ("0", "1"):
        monkeypatch.setenv("MSWB_MT_JUMP", mode)
        merged = np.full_like(whole, np.nan)
        for r in range(W):
            t, it = lik.bootstrap_run(B, seed=-5, algo=algo, replica_rank=r, replica_world=W)
            mine = np.arange(B) % W == r
            assert np.all(np.isnan(t[~mine])) and not np.any(np.isnan(t[mine]))
            merged[mine] = t[mine]
            assert [it[i] for i in np.flatnonzero(mine)] == [it_whole[i] for i in np.flatnonzero(mine)]
        assert np.array_equal(merged, whole), mode
    # a fixed number of draws per replicate (--bootstrap-count) moves the replicate boundaries
    monkeypatch.setenv("MSWB_MT_JUMP", "1")
    whole, _ = lik.bootstrap_run(5, seed=11, algo=algo, bootstrap_count=1234)
    t, _ = lik.bootstrap_run(5, seed=11, algo=algo, bootstrap_count=1234, replica_rank=1, replica_world=2)
    assert np.array_equal(t[[1, 3]], whole[[1, 3]])


@pytest.mark.parametrize("storage,K", [("f64", 6), ("f64", 300), ("f64", 1100), ("f32", 300)])
def test_bootstrap_em_runs_as_device_batches(oracle, mswb, ctx, storage, K):
    """north_star (3): with EM on a dense likelihood the replicates of a GPU run as ONE batch — every count vector is
    resampled first, then each sweep of the matrix serves all replicates still running (em_lin_batch_kernel).  Every
    replicate must equal the reference loop (src/mSWEEP.cpp:496-518): its own cold start, its own stopping iteration."""
    S = 4 if K > 100 else 20
    wl = synth.generate_ec_patterns(6000, K, S, n_present=min(5, K), seed=40 + K, dup_factor=2.0)
    ec = oracle.ec_build_csr(wl.n_reads, wl.n_targets, wl.row_ptr, wl.targets)
    ref_l = oracle.lik_build(ec, wl.group_of_target, wl.group_sizes)
    aln = mswb.Alignment(ctx, wl.n_reads, wl.n_targets, wl.row_ptr, wl.targets)
    f32 = storage == "f32"
    lik = mswb.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes, storage=mswb.STORE_F32 if f32 else mswb.STORE_F64)
    B = 7                                   # not a multiple of the replicates per CTA: the last slice is ragged
    tol, cap = (0.0, 25) if f32 else (1e-6, 3000)      # fp32 storage: compared at a fixed iteration count
    l0 = mswb.launch_count()
    thetas, iters = lik.bootstrap_run(B, seed=7, algo=mswb.ALGO_EM, tol=tol, max_iters=cap)
    launches_batched = mswb.launch_count() - l0
    counts = oracle.bootstrap_resample(ec.count, 7, B)
    for r in range(B):
        with np.errstate(divide="ignore"):
            ref = oracle.vi_run("em", ref_l.logl, np.log(counts[r].astype(np.float64)), tol=tol, max_iters=cap)
        assert iters[r] == ref.iters, (r, iters[r], ref.iters)
        assert np.max(np.abs(thetas[r] - ref.theta)) < (2e-6 if f32 else THETA_TOL)
    # the one-by-one path (MSWB_BOOT_BATCH=0) gives the same answers with several times the launches
    os.environ["MSWB_BOOT_BATCH"] = "0"
    try:
        l0 = mswb.launch_count()
        seq, seq_iters = lik.bootstrap_run(B, seed=7, algo=mswb.ALGO_EM, tol=tol, max_iters=cap)
        launches_seq = mswb.launch_count() - l0
    finally:
        del os.environ["MSWB_BOOT_BATCH"]
    assert seq_iters == iters
    assert np.max(np.abs(seq - thetas)) < (1e-6 if f32 else 1e-10)
    assert launches_batched < launches_seq
    # replicas spread over two ranks: each batches its own rows, the union equals the single run
    t0, _ = lik.bootstrap_run(B, seed=7, algo=mswb.ALGO_EM, tol=tol, max_iters=cap, replica_rank=0, replica_world=2)
    t1, _ = lik.bootstrap_run(B, seed=7, algo=mswb.ALGO_EM, tol=tol, max_iters=cap, replica_rank=1, replica_world=2)
    assert np.array_equal(np.isnan(t0), ~np.isnan(t1))
    assert np.max(np.abs(np.where(np.isnan(t0), t1, t0) - thetas)) < 1e-12


def test_trim_and_abort_on_a_single_gpu_context(mswb):
    """mswb_ctx_trim hands the parked blocks back; mswb_ctx_abort without a communicator is a no-op; the context stays usable."""
    c = mswb.Context(0)
    wl = synth.generate(3000, 90, 6, n_present=2, n_templates=30, seed=8)
    aln = mswb.Alignment(c, wl.n_reads, wl.n_targets, wl.row_ptr, wl.targets)
    lik = mswb.Likelihood.build(c, aln, wl.group_of_target, wl.group_sizes)
    first = lik.vi_run(mswb.ALGO_EM)
    c.trim()
    c.abort()
    again = lik.vi_run(mswb.ALGO_EM)
    assert again.iters == first.iters and np.array_equal(again.theta, first.theta)
    lik.close(); aln.close(); c.close()
