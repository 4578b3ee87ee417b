"""An INDEPENDENT restatement of the optimiser of the path (rcgpar's RCG and EM/VB for the mSWEEP mixture model),
written from the model's equations and the recollection in SURVEY.md §8(a) only — not from oracle/oracle.cpp, whose
structure (group-major loops, literal revert-and-resubtract restart, double accumulators) it deliberately does not
share.  numpy, class-major arrays, extended precision (np.longdouble) for every sum over classes.

TEST INFRASTRUCTURE.  tests/golden/make_vi_independent.py freezes its trajectories in tests/golden/vi_independent.npz;
tests/test_vi_independent.py holds the oracle (CPU) and the CUDA path (GPU) against them.  It does NOT pin rcgpar
itself (off-tree, v1.2.1): both restatements start from the same recollection of the algorithm.  What it does catch is
a transcription error in either one — they only agree if both implement the equations below.

Model (Mäklin et al. 2021; the BitSeqVB family): reads of equivalence class j (c_j of them) come from group k with
probability theta_k and likelihood exp(logl[k, j]); theta ~ Dirichlet(alpha0).  Mean field q(theta) q(z) with
q(z_j = k) = phi_jk = exp(gamma_jk).  With q(theta) at its optimum given phi — Dirichlet(N), N_k = alpha0_k + sum_j c_j phi_jk —
the bound is a function of phi alone:

    L(phi) = sum_j c_j sum_k phi_jk (logl_kj - log phi_jk) + sum_k lgamma(N_k) - lgamma(sum_k N_k)
             + lgamma(sum_k alpha0_k) - sum_k lgamma(alpha0_k),          sum_k N_k = sum_k alpha0_k + sum_j c_j.

EM / VB: phi_jk  proportional to  exp(logl_kj + digamma(N_k)).
RCG: in the softmax coordinates gamma the natural-gradient direction is d_jk = logl_kj + digamma(N_k) - 1 - gamma_jk; its squared
Riemannian norm is sum_jk phi_jk (d_jk - <d>_j) d_jk with <d>_j = sum_k phi_jk d_jk; Fletcher-Reeves conjugation
beta = |g_new|^2 / |g_old|^2, step = d + beta * previous step; gamma += step, each class renormalised.  A step that
lowers the bound is replaced by the plain step from the same point (which, renormalised, is
gamma = normalise(logl + digamma(N_k))) and the direction memory is dropped.
digamma is the authors' approximation (src/Sample.cpp:87-97 of the reference): recurrence up to 7, then a series in 1/(x - 1/2).
"""
from __future__ import annotations

import numpy as np
from scipy.special import gammaln

LD = np.longdouble


def digamma_authors(x):
    """src/Sample.cpp:87-97, vectorised: psi(x) = psi(x + n) - sum 1/(x + i) until x >= 7, then
    log(y) + 1/(24 y^2) - 7/(960 y^4) + 31/(8064 y^6) - 127/(30720 y^8) with y = x - 1/2."""
    x = np.array(x, dtype=np.float64, copy=True)
    shift = np.zeros_like(x)
    while True:
        low = x < 7.0
        if not low.any():
            break
        shift[low] -= 1.0 / x[low]
        x[low] += 1.0
    y = x - 0.5
    y2 = 1.0 / (y * y)
    y4 = y2 * y2
    return shift + np.log(y) + y2 / 24.0 - 7.0 / 960.0 * y4 + 31.0 / 8064.0 * y4 * y2 - 127.0 / 30720.0 * y4 * y4


def _normalise(g):
    """log-softmax over the groups of every class; g is (N, K)."""
    m = g.max(axis=1, keepdims=True)
    return g - (m + np.log(np.exp(g - m).sum(axis=1, keepdims=True)))


def _counts(log_counts):
    with np.errstate(under="ignore"):
        return np.exp(np.asarray(log_counts, np.float64))        # exp(-inf) = 0: a class that was not resampled


def _expected_counts(gamma, c, alpha0):
    phi = np.exp(gamma)
    return alpha0 + np.asarray((phi.astype(LD) * c.astype(LD)[:, None]).sum(axis=0), np.float64)


def _bound(gamma, logl_t, c, N_k, alpha0):
    phi = np.exp(gamma)
    w = phi * c[:, None]
    term = np.where(w > 0, w * (logl_t - gamma), 0.0)           # a zero-count class contributes nothing
    data = term.astype(LD).sum()
    const = gammaln(alpha0.sum()) - gammaln(alpha0.sum() + c.astype(LD).sum().astype(np.float64)) - gammaln(alpha0).astype(LD).sum()
    return float(data + gammaln(N_k).astype(LD).sum() + const)


def run(algo: str, logl, log_counts, alpha0=None, tol=1e-6, max_iters=5000):
    """logl: (K, N) group-major as the reference holds it.  Returns dict(theta, N_k, bound, iters, converged, and the
    per-iteration traces bound / gnorm / reset)."""
    logl_t = np.ascontiguousarray(np.asarray(logl, np.float64).T)        # class-major (N, K)
    N, K = logl_t.shape
    alpha0 = np.ones(K) if alpha0 is None else np.asarray(alpha0, np.float64)
    c = _counts(log_counts)
    gamma = np.full((N, K), np.log(1.0 / K))
    N_k = _expected_counts(gamma, c, alpha0)
    tb, tg, tr = [], [], []
    converged = False
    if algo == "em":
        bound = 0.0
        for it in range(max_iters):
            gamma = _normalise(logl_t + digamma_authors(N_k)[None, :])
            N_k = _expected_counts(gamma, c, alpha0)
            old, bound = bound, _bound(gamma, logl_t, c, N_k, alpha0)
            tb.append(bound); tg.append(0.0); tr.append(0)
            if it > 0 and abs(bound - old) < tol:
                converged = True
                break
    elif algo == "rcg":
        bound, old_norm, had_reset = -100000.0, 1.0, False
        prev_step = np.zeros((N, K))
        for it in range(max_iters):
            psi = digamma_authors(N_k) - 1.0
            d = logl_t + psi[None, :] - gamma
            phi = np.exp(gamma)
            mean_d = (phi * d).sum(axis=1, keepdims=True)
            new_norm = float((phi * (d - mean_d) * d).astype(LD).sum())
            beta = new_norm / old_norm
            old_norm = new_norm
            step = d.copy()
            if not had_reset and beta > 0:
                step += beta * prev_step
            had_reset = False
            cand = _normalise(gamma + step)
            cand_N = _expected_counts(cand, c, alpha0)
            cand_bound = _bound(cand, logl_t, c, cand_N, alpha0)
            old = bound
            if cand_bound < old:
                # the conjugate direction lost ground: the plain step from the same point, direction memory dropped
                had_reset = True
                gamma = _normalise(logl_t + psi[None, :])
                N_k = _expected_counts(gamma, c, alpha0)
                bound = _bound(gamma, logl_t, c, N_k, alpha0)
                # (the memory is not advanced: the next iteration ignores it anyway and overwrites it)
            else:
                gamma, N_k, bound = cand, cand_N, cand_bound
                prev_step = step
            tb.append(bound); tg.append(new_norm); tr.append(1 if had_reset else 0)
            if bound - old < tol and not had_reset:
                converged = True
                break
    else:
        raise ValueError(algo)
    theta = np.asarray((np.exp(gamma).astype(LD) * c.astype(LD)[:, None]).sum(axis=0) / c.astype(LD).sum(), np.float64)
    return {"theta": theta, "N_k": N_k, "bound": bound, "iters": len(tb), "converged": converged,
            "trace_bound": np.array(tb), "trace_gnorm": np.array(tg), "trace_reset": np.array(tr, np.uint8), "gamma": gamma.T.copy()}
