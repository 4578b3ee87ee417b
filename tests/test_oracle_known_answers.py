"""CPU: the oracle against known answers that follow from the reference's in-tree code alone
(computed independently with Python integers / scipy in tests/golden/make_golden.py) and against the
part of the real reference that compiles here (oracle/_ref: src/Grouping.cpp + src/Reference.cpp)."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def ka():
    return np.load(os.path.join(GOLD, "known_answers.npz"))


def test_hash_fold_single_target(oracle):
    # include/mSWEEP_alignment.hpp:150-155 with hash = 0, j = 0
    assert oracle.pattern_hash([0]) == 0x517cc1b727220a95
    assert oracle.pattern_hash([]) == 0


def test_hash_fold_vectors(oracle, ka):
    ptr, tg = ka["hash_ptr"].astype(np.int64), ka["hash_targets"]
    for i, want in enumerate(ka["hashes"]):
        assert oracle.pattern_hash(tg[ptr[i]:ptr[i + 1]]) == int(want)


def test_hash_is_order_dependent(oracle):
    assert oracle.pattern_hash([1, 2]) != oracle.pattern_hash([2, 1])


def test_lut_against_scipy(oracle, ka):
    # include/Likelihood.hpp:47-60, 92-107: LUT[0] = log(zi), LUT[c] = ldbb_scaled(c, n, a, b) + log1p(-zi)
    for n, (a_ref, b_ref) in zip(ka["lut_sizes"], ka["bb_alpha_beta"]):
        a, b = oracle.bb_parameters(int(n), 0.65, 0.01)
        assert a == pytest.approx(a_ref, rel=1e-15) and b == pytest.approx(b_ref, rel=1e-15)
        ref = ka[f"lut_{n}"]
        got = np.array([np.log(0.01)] + [oracle.ldbb_scaled(c, int(n), a, b) + np.log1p(-0.01) for c in range(1, int(n) + 1)])
        assert np.max(np.abs(got - ref)) < 1e-11 * max(1.0, np.max(np.abs(ref)))
        assert got[-1] == pytest.approx(np.log1p(-0.01), abs=1e-12)      # ldbb_scaled(n, n, ., .) == 0


def test_lut_spot_values(oracle):
    # n = 60, q = 0.65, e = 0.01
    a, b = oracle.bb_parameters(60, 0.65, 0.01)
    assert a == pytest.approx(1.856258924322, abs=1e-11)
    assert b == pytest.approx(0.999524036173, abs=1e-11)
    for n, want in [(2, -0.3717), (7, -1.2199), (15, -1.8167), (60, -2.9666), (255, -4.1971), (1000, -5.3662)]:
        a, b = oracle.bb_parameters(n, 0.65, 0.01)
        assert oracle.ldbb_scaled(1, n, a, b) + np.log1p(-0.01) == pytest.approx(want, abs=6e-5)


def test_digamma_series(oracle, ka):
    got = np.array([oracle.digamma(float(x)) for x in ka["digamma_x"]])
    assert np.max(np.abs(got - ka["digamma_ref"])) < 1e-10     # measured 5.3e-11, worst at the x ~ 6-7 hand-over


def _write(tmp_path, lines, name="grouping.txt"):
    p = tmp_path / name
    p.write_text("".join(l + "\n" for l in lines))
    return str(p)


@pytest.mark.parametrize("n_names,n_lines,seed", [(3, 10, 0), (40, 500, 1), (300, 2000, 2), (70000, 70500, 3)])
def test_grouping_matches_real_reference(oracle, tmp_path, n_names, n_lines, seed):
    """oracle.read_grouping == the reference's own Reference/Grouping classes compiled from its sources."""
    if not oracle.ref_grouping_available():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    rng = np.random.default_rng(seed)
    names = [f"lin_{i}" for i in rng.permutation(n_names)]
    lines = [names[i] for i in rng.integers(0, n_names, size=n_lines)]
    lines[:0] = names[: min(n_names, 50)]
    path = _write(tmp_path, lines)
    n_o, s_o, g_o = oracle.read_grouping(path)
    n_r, s_r, g_r = oracle.ref_read_grouping(path)
    assert n_o == n_r
    assert np.array_equal(s_o, s_r)
    assert np.array_equal(g_o, g_r)
    assert g_o[0] == 0 and n_o[0] == lines[0]          # ids in order of first appearance


def test_collapse_semantics(oracle):
    """include/mSWEEP_alignment.hpp:137-215 on a hand-made table."""
    rows = [[3, 5], [], [1], [3, 5], [0, 9], [1], [3, 5], []]
    ptr = np.cumsum([0] + [len(r) for r in rows]).astype(np.uint64)
    tg = np.array([t for r in rows for t in r], np.uint32)
    ec = oracle.ec_build_csr(len(rows), 10, ptr, tg)
    want = sorted({oracle.pattern_hash(r) for r in rows if r})
    assert ec.n_reads == 8
    assert list(ec.hash) == want                                    # ascending unsigned hash (std::map order)
    by_hash = {int(h): i for i, h in enumerate(ec.hash)}
    i35, i1, i09 = by_hash[oracle.pattern_hash([3, 5])], by_hash[oracle.pattern_hash([1])], by_hash[oracle.pattern_hash([0, 9])]
    assert list(ec.count[[i35, i1, i09]]) == [3, 2, 1]
    assert list(ec.rep_read[[i35, i1, i09]]) == [0, 2, 4]           # smallest read id of each class
    rp = ec.read_ptr.astype(int)
    assert list(ec.read_ids[rp[i35]:rp[i35 + 1]]) == [0, 3, 6]      # ascending inside a class
    pp = ec.pat_ptr.astype(int)
    assert list(ec.pat_targets[pp[i09]:pp[i09 + 1]]) == [0, 9]
    assert int(ec.count.sum()) == 6                                 # unaligned reads are in no class


def test_likelihood_hand_case(oracle):
    """include/Likelihood.hpp:109-186 on a hand-made table incl. --min-hits."""
    rows = [[0, 1, 4], [0, 1, 4], [2], [5]]
    ptr = np.cumsum([0] + [len(r) for r in rows]).astype(np.uint64)
    tg = np.array([t for r in rows for t in r], np.uint32)
    ec = oracle.ec_build_csr(4, 8, ptr, tg)
    got = np.array([0, 0, 1, 1, 2, 2, 3, 3], np.uint32)             # 4 groups of 2 targets
    sizes = np.array([2, 2, 2, 2], np.uint64)
    L = oracle.lik_build(ec, got, sizes, keep_hit_counts=True)
    by_hash = {int(h): i for i, h in enumerate(ec.hash)}
    i014 = by_hash[oracle.pattern_hash([0, 1, 4])]
    assert list(L.hit_counts[:, i014]) == [2, 0, 1, 0]
    a, b = oracle.bb_parameters(2, 0.65, 0.01)
    assert L.logl[0, i014] == oracle.ldbb_scaled(2, 2, a, b) + np.log1p(-0.01)
    assert L.logl[1, i014] == np.log(0.01)
    assert L.log_counts[i014] == np.log(2.0)
    # min_hits = 2: group 0 sees 2 reads, group 2 sees 2 + 1, group 1 sees 1, group 3 none
    L2 = oracle.lik_build(ec, got, sizes, min_hits=2)
    assert list(L2.hits) == [2, 1, 3, 0]
    assert list(L2.mask) == [1, 0, 1, 0]
    assert L2.n_groups == 2 and L2.logl.shape == (2, 3)
    assert np.array_equal(L2.logl, L.logl[[0, 2]])


def test_bootstrap_draws(oracle):
    """src/BootstrapSample.cpp:60-73: counts sum to bootstrap_count, same seed same stream, replicates
    are consecutive draws of ONE generator, and the restated libstdc++ recipe reproduces the counts."""
    counts = np.array([5, 1, 0, 30, 7, 2], np.uint64)
    a = oracle.bootstrap_resample(counts, seed=7, n_replicates=3)
    b = oracle.bootstrap_resample(counts, seed=7, n_replicates=3)
    assert np.array_equal(a, b)
    assert list(a.sum(axis=1)) == [45, 45, 45]
    assert np.all(a[:, 2] == 0)
    c = oracle.bootstrap_resample(counts, seed=7, n_replicates=1, bootstrap_count=1000)
    assert int(c.sum()) == 1000
    # independent restatement: mt19937_64 in Python integers + lower_bound on the cumulative table
    from tests.mt64 import MT19937_64
    gen = MT19937_64(7)
    p = counts.astype(np.float64) / float(counts.astype(np.float64).sum())
    cp = np.cumsum(p)
    cp[-1] = 1.0
    want = np.zeros((3, len(counts)), np.uint32)
    for r in range(3):
        for _ in range(45):
            u = float(gen.next()) * 2.0 ** -64
            if u >= 1.0:
                u = np.nextafter(1.0, 0.0)
            want[r, int(np.searchsorted(cp, u, side="left"))] += 1
    assert np.array_equal(a, want)


def test_negative_seed_sign_extends(oracle):
    counts = np.array([3, 3, 3], np.uint64)
    from tests.mt64 import MT19937_64
    gen = MT19937_64((-7) & ((1 << 64) - 1))
    got = oracle.bootstrap_resample(counts, seed=-7, n_replicates=1)[0]
    cp = np.array([1 / 3, 1 / 3 + 1 / 3, 1.0])
    want = np.zeros(3, np.uint32)
    for _ in range(9):
        want[int(np.searchsorted(cp, float(gen.next()) * 2.0 ** -64, side="left"))] += 1
    assert np.array_equal(got, want)
