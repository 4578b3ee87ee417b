"""Regenerates tests/golden/*.npz.

The reference ships no tests, fixtures or example data, so there is nothing of its own to
pin against beyond the in-tree formulas.  These fixtures freeze
  (a) known answers that follow from the reference's in-tree code alone, computed INDEPENDENTLY of the
      oracle with scipy / plain Python integers (hash fold, LL_WOR21 lookup table, beta-binomial
      parameters, the authors' digamma series), and
  (b) the oracle's outputs on one small seeded pseudoalignment (EC table, hit counts, --min-hits mask,
      log-likelihoods, RCG / EM runs, bootstrap counts), so that the GPU box — where /root/reference
      does not exist — checks the CUDA path against frozen numbers as well as against the live oracle.

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
from scipy.special import gammaln, digamma

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from msweep_b200 import synth            # noqa: E402
from oracle import pyoracle as orc       # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
M64 = (1 << 64) - 1


def py_hash(targets):
    """include/mSWEEP_alignment.hpp:150-155 in Python integers."""
    h = 0
    for j in targets:
        h ^= (j + 0x517cc1b727220a95 + ((h << 6) & M64) + (h >> 2)) & M64
    return h


def scipy_lut(n, q=0.65, e=0.01, zi=0.01):
    """include/Likelihood.hpp:47-60, 92-107, 198-207 with scipy.special.gammaln."""
    mean = n * q
    phi = 1.0 / (n - mean + e)
    beta = phi * (n - mean)
    alpha = mean * beta / (n - mean)
    lbeta = lambda x, y: gammaln(x) + gammaln(y) - gammaln(x + y)
    k = np.arange(1, n + 1, dtype=np.float64)
    v = gammaln(n + 1) - gammaln(k + 1) - gammaln(n - k + 1) + lbeta(k + alpha, n - k + beta) - lbeta(n + alpha, beta)
    return np.concatenate(([np.log(zi)], v + np.log1p(-zi))), alpha, beta


def main():
    rng = np.random.default_rng(11)
    # (a) known answers ---------------------------------------------------------------------------
    pats = [[0], [5], [0, 1, 2], [3, 17, 59, 1000, 2999], list(range(0, 3000, 7))]
    pats += [sorted(rng.choice(60000, size=int(n), replace=False).tolist()) for n in rng.integers(1, 80, size=40)]
    flat = np.array([t for p in pats for t in p], np.uint32)
    ptr = np.cumsum([0] + [len(p) for p in pats]).astype(np.uint64)
    hashes = np.array([py_hash(p) for p in pats], np.uint64)
    sizes = [1, 2, 7, 15, 60, 255, 1000]
    luts = {f"lut_{n}": scipy_lut(n)[0] for n in sizes}
    ab = np.array([scipy_lut(n)[1:] for n in sizes])
    xs = np.concatenate([np.logspace(-3, 9, 400), np.linspace(0.5, 12, 200)])
    np.savez(os.path.join(OUT, "known_answers.npz"), hash_ptr=ptr, hash_targets=flat, hashes=hashes,
             lut_sizes=np.array(sizes), bb_alpha_beta=ab, digamma_x=xs, digamma_ref=digamma(xs), **luts)

    # (b) frozen oracle outputs on a small seeded case ---------------------------------------------
    wl = synth.generate(4000, 120, 8, n_present=3, n_templates=60, p_noise=0.05, seed=20231017)
    ec = orc.ec_build_csr(wl.n_reads, wl.n_targets, wl.row_ptr, wl.targets)
    lik = orc.lik_build(ec, wl.group_of_target, wl.group_sizes, keep_hit_counts=True)
    lik_mh = orc.lik_build(ec, wl.group_of_target, wl.group_sizes, min_hits=1)
    rcg = orc.vi_run("rcg", lik.logl, lik.log_counts, tol=1e-6, max_iters=5000)
    em = orc.vi_run("em", lik.logl, lik.log_counts, tol=1e-6, max_iters=5000)
    boot = orc.bootstrap_resample(ec.count, seed=42, n_replicates=3)
    np.savez_compressed(
        os.path.join(OUT, "small_case.npz"),
        n_reads=wl.n_reads, n_targets=wl.n_targets, row_ptr=wl.row_ptr, targets=wl.targets,
        group_of_target=wl.group_of_target, group_sizes=wl.group_sizes,
        ec_hash=ec.hash, ec_count=ec.count, ec_rep_read=ec.rep_read, ec_pat_ptr=ec.pat_ptr, ec_pat_targets=ec.pat_targets,
        ec_read_ptr=ec.read_ptr, ec_read_ids=ec.read_ids,
        hit_counts=lik.hit_counts, logl=lik.logl, log_counts=lik.log_counts,
        mask_minhits1=lik_mh.mask, hits_minhits1=lik_mh.hits,
        rcg_theta=rcg.theta, rcg_bound=rcg.bound, rcg_iters=rcg.iters, rcg_trace=rcg.trace_bound,
        em_theta=em.theta, em_bound=em.bound, em_iters=em.iters,
        boot_seed=42, boot_counts=boot)
    print("ECs:", ec.n_ecs, "rcg iters:", rcg.iters, "em iters:", em.iters)


if __name__ == "__main__":
    main()
