"""GPU: size-independent properties at sizes where the oracle would take minutes (BASELINE configs 2-4 shapes):
sortedness / conservation / idempotence of the EC build, conservation of the likelihood build, and
model invariants of the optimiser on a multi-GB matrix."""
import numpy as np
import pytest

from msweep_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big(mswb, ctx):
    wl = synth.generate_ec_patterns(1_500_000, 1000, 24, n_present=40, seed=77, dup_factor=1.5)
    aln = mswb.Alignment(ctx, wl.n_reads, wl.n_targets, wl.row_ptr, wl.targets)
    return wl, aln


def test_ec_build_properties_at_scale(big, mswb, ctx):
    wl, aln = big
    e = aln.export()
    lens = np.diff(wl.row_ptr.astype(np.int64))
    assert np.all(np.diff(e.hash.astype(np.uint64)) > 0)                     # strictly ascending: std::map order, no duplicate keys
    assert int(e.count.sum()) == int((lens > 0).sum()) == aln.n_aligned        # every aligned read in exactly one class
    assert np.array_equal(np.diff(e.read_ptr.astype(np.int64)), e.count.astype(np.int64))
    assert np.array_equal(np.sort(e.read_ids), np.nonzero(lens > 0)[0].astype(np.uint32))   # a permutation of the aligned reads
    first = e.read_ids[e.read_ptr[:-1].astype(np.int64)]
    assert np.array_equal(first, e.rep_read)                                  # representative = first (smallest) member
    seg_min = np.minimum.reduceat(e.read_ids, e.read_ptr[:-1].astype(np.int64))
    assert np.array_equal(seg_min, e.rep_read)
    assert np.array_equal(np.diff(e.pat_ptr.astype(np.int64)), lens[e.rep_read])
    # idempotence: collapsing the class patterns again gives the same classes, each seen once
    again = mswb.Alignment(ctx, aln.n_ecs, wl.n_targets, e.pat_ptr, e.pat_targets).export()
    assert np.array_equal(again.hash, e.hash) and np.all(again.count == 1)
    assert np.array_equal(again.pat_targets, e.pat_targets)
    # spot-check hashes against the host fold (mswb_pattern_hash)
    pp = e.pat_ptr.astype(np.int64)
    for i in np.linspace(0, aln.n_ecs - 1, 50).astype(int):
        assert mswb.pattern_hash(e.pat_targets[pp[i]:pp[i + 1]]) == int(e.hash[i])


def test_likelihood_and_optimiser_invariants_at_scale(big, mswb, ctx):
    wl, aln = big
    lik = mswb.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes)             # ~1e6 x 1000 fp64 = 8 GB
    # every lineage gets a few stray hits in this workload: prune the half of the lineages with the fewest
    _, hits0 = mswb.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes, min_hits=1).mask(want_hits=True)
    thr = int(np.sort(hits0)[wl.n_groups // 2])
    lik1 = mswb.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes, min_hits=thr)
    mask, hits = lik1.mask(want_hits=True)
    # tallies by hand from the exported class table
    e = aln.export()
    grp = wl.group_of_target[e.pat_targets]
    cls = np.repeat(np.arange(aln.n_ecs), np.diff(e.pat_ptr.astype(np.int64)))
    pairs = np.unique(np.stack([cls, grp.astype(np.int64)]), axis=1)
    want = np.bincount(pairs[1], weights=e.count[pairs[0]].astype(np.float64), minlength=wl.n_groups).astype(np.uint64)
    assert np.array_equal(hits, want) and np.array_equal(mask, (want >= thr).astype(np.uint8))
    assert lik1.n_groups == int(mask.sum()) < wl.n_groups

    em = lik.vi_begin(mswb.ALGO_EM, tol=0.0, max_iters=40)
    em.step(40)
    tb, _, _ = em.trace()
    r = em.finish()
    assert np.all(np.diff(tb) >= -1e-9 * np.abs(tb[:-1]))                     # EM never decreases the bound
    assert abs(r.theta.sum() - 1.0) < 1e-12
    total = float(aln.n_aligned)
    assert np.max(np.abs((r.N_k - 1.0) / total - r.theta)) < 1e-15            # theta = (N_k - alpha0) / sum c

    rcg = lik.vi_run(mswb.ALGO_RCG, tol=1e-6, max_iters=400)
    assert rcg.converged and abs(rcg.theta.sum() - 1.0) < 1e-12
    em_long = lik.vi_run(mswb.ALGO_EM, tol=1e-7, max_iters=3000)
    assert np.max(np.abs(em_long.theta - rcg.theta)) < 2e-4                   # same optimum, different paths
    assert abs(em_long.bound - rcg.bound) < 1e-6 * abs(rcg.bound)
    # pruning lineages with few hits moves the estimates of the others by about the pruned mass (README.md:136-140)
    r1 = lik1.vi_run(mswb.ALGO_RCG, tol=1e-6, max_iters=400)
    pruned_mass = float(rcg.theta[~mask.astype(bool)].sum())
    assert pruned_mass < 0.05
    assert np.max(np.abs(r1.theta - rcg.theta[mask.astype(bool)])) < max(5e-4, 2 * pruned_mass)
    # fp32 storage against fp64 at a fixed iteration count
    lik32 = mswb.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes, storage=mswb.STORE_F32)
    a = lik32.vi_run(mswb.ALGO_EM, tol=0.0, max_iters=30)
    b = lik.vi_run(mswb.ALGO_EM, tol=0.0, max_iters=30)
    assert np.max(np.abs(a.theta - b.theta)) < 2e-6 and abs(a.bound - b.bound) < 1e-6 * abs(b.bound)
    # run-to-run bit reproducibility (no atomics on the path)
    c = lik.vi_run(mswb.ALGO_RCG, tol=1e-6, max_iters=400)
    assert np.array_equal(c.theta, rcg.theta) and c.bound == rcg.bound and c.iters == rcg.iters


def test_sparse_storage_at_scale(big, mswb, ctx):
    """Both optimisers on the sparse storage against the dense fp64 sweeps on the same ~1e6 x 1000 problem: RCG iteration by
    iteration over a fixed number of iterations (before the stopping rule — decided at a relative 1e-14 of the bound here —
    gets a say), the converged abundances, bit-reproducibility, posterior tiles."""
    wl, aln = big
    dense = mswb.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes)
    sparse = mswb.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes, storage=mswb.STORE_SPARSE)
    traces = {}
    for name, lik in (("dense", dense), ("sparse", sparse)):
        s = lik.vi_begin(mswb.ALGO_RCG, tol=-1e300, max_iters=60)
        s.step(60)
        tb, tg, tr = s.trace()
        traces[name] = (tb, tg, tr, s.finish())
    (tb_d, tg_d, tr_d, r_d), (tb_s, tg_s, tr_s, r_s) = traces["dense"], traces["sparse"]
    # the same restart pattern and bound, iteration by iteration — as long as the accept / reject decisions are not
    # themselves decided by rounding (late in the run a step may change the bound by less than one ulp of 1e7)
    differ = np.flatnonzero(tr_d != tr_s)
    n_same = int(differ[0]) if len(differ) else len(tr_d)
    assert n_same >= 30, (n_same, tr_d, tr_s)
    assert np.max(np.abs(tb_d[:n_same] - tb_s[:n_same]) / np.abs(tb_d[:n_same])) < 1e-11
    assert np.allclose(tg_d[:30], tg_s[:30], rtol=1e-6, atol=1e-9 * tg_d[0])
    if n_same == len(tr_d):
        assert np.max(np.abs(r_d.theta - r_s.theta)) < 1e-9
    conv_d, conv_s = dense.vi_run(mswb.ALGO_RCG), sparse.vi_run(mswb.ALGO_RCG)
    assert conv_d.converged and conv_s.converged and abs(conv_d.iters - conv_s.iters) <= 3
    assert np.max(np.abs(conv_d.theta - conv_s.theta)) < 1e-6 and abs(conv_d.bound - conv_s.bound) < 1e-9 * abs(conv_d.bound)
    again = sparse.vi_run(mswb.ALGO_RCG)
    assert np.array_equal(again.theta, conv_s.theta) and again.bound == conv_s.bound and again.iters == conv_s.iters
    # posteriors of the two converged runs: they may stop up to 3 iterations apart (asserted above; the stopping rule is
    # decided by the summation order at this size), and a class's responsibilities move more than theta does in an iteration
    g_d, g_s = dense.posteriors(1000, 1300), sparse.posteriors(1000, 1300)
    assert np.max(np.abs(np.exp(g_d) - np.exp(g_s))) < (1e-7 if conv_d.iters == conv_s.iters else 1e-5)
    em_d, em_s = dense.vi_run(mswb.ALGO_EM, tol=0.0, max_iters=50), sparse.vi_run(mswb.ALGO_EM, tol=0.0, max_iters=50)
    assert np.max(np.abs(em_d.theta - em_s.theta)) < 1e-11 and abs(em_d.bound - em_s.bound) < 1e-12 * abs(em_d.bound)


@pytest.mark.parametrize("algo_name", ["rcg", "em"])
def test_fused_iterations_at_scale_equal_the_launch_per_sweep_path(big, mswb, ctx, monkeypatch, algo_name):
    """~1e6 classes x 1000 groups (config 2 / 5 size): whole iterations inside one cooperative launch (rcgs_fused_kernel /
    ems_fused_kernel with the partial vectors reduced by every CTA on its tiles of columns between two grid rendezvous)
    against one launch per sweep + reduction kernel.  Same sweeps, same order over the partial vectors; the K-sized sums of
    the control step (lgamma terms of the bound) run over 256 instead of 1024 threads, so the trajectories agree to the last
    digits of the bound rather than bit for bit."""
    wl, aln = big
    sparse = mswb.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes, storage=mswb.STORE_SPARSE)
    algo = mswb.ALGO_RCG if algo_name == "rcg" else mswb.ALGO_EM
    runs = {}
    for fused in ("0", "1"):
        monkeypatch.setenv("MSWB_FUSED", fused)
        n0 = mswb.launch_count()
        s = sparse.vi_begin(algo, tol=-1e300 if algo_name == "rcg" else 0.0, max_iters=40)
        s.step(25); s.poll(); s.step(15)
        tb, tg, tr = s.trace()
        runs[fused] = (tb, tg, tr, s.finish(), mswb.launch_count() - n0)
    (tb0, tg0, tr0, r0, n_launch), (tb1, tg1, tr1, r1, n_fused) = runs["0"], runs["1"]
    assert r0.iters == r1.iters == 40 and r0.resets == r1.resets and np.array_equal(tr0, tr1)
    assert np.max(np.abs(tb0 - tb1) / np.abs(tb0)) < 1e-13
    assert np.allclose(tg0, tg1, rtol=1e-9, atol=1e-12 * max(1.0, float(np.max(np.abs(tg0)))))
    assert np.max(np.abs(r0.theta - r1.theta)) < 1e-12 and abs(r0.bound - r1.bound) < 1e-13 * abs(r0.bound)
    assert n_fused < n_launch / 5, (n_fused, n_launch)


def test_staged_host_to_device_copy_builds_the_same_table():
    """Large pageable inputs cross PCIe through several host threads with pinned bounce buffers (ctx.cu: h2d_staged); the
    switches are read once per process, so the two variants run as two processes (tools/check_h2d_staged.py)."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = {}
    for tag, env in (("staged", {"MSWB_H2D_STAGED_MIN_MB": "1"}), ("plain", {"MSWB_H2D_STAGED": "0"})):
        r = subprocess.run([sys.executable, os.path.join(root, "tools", "check_h2d_staged.py")], capture_output=True, text=True, timeout=600,
                           env=dict(os.environ, **env))
        assert r.returncode == 0, r.stderr[-2000:]
        out[tag] = [l for l in r.stdout.splitlines() if l.startswith("digest")][-1].split()[1]
    assert out["staged"] == out["plain"]
