"""Regenerates tests/golden/vi_independent.npz: trajectories of the INDEPENDENT numpy restatement of the optimiser
(tests/ref_numpy_vi.py) on a few seeded dense problems.  The inputs are regenerated from the seeds by the tests.

Run from the repo root:  python tests/golden/make_vi_independent.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import ref_numpy_vi as ref   # noqa: E402

# name: (K, N, seed) — Gaussian log-likelihoods — or (K, N, seed, count scale) — LL_WOR21-like rows: log(0.01) everywhere but
# on up to three hit groups, large class counts; the first RCG step falls below the starting bound of -1e5 there, so
# these cases go through the restart branch at iteration 0.
CASES = {"k7": (7, 300, 11), "k40": (40, 1500, 12), "k3_zero_counts": (3, 50, 13), "k130": (130, 900, 14),
         "wor50_restart": (50, 2000, 20, 50), "wor20_restart": (20, 3000, 20, 1000)}
TOL = 1e-8
TOL_BY_CASE = {"wor50_restart": 1e-6, "wor20_restart": 1e-4}   # (bounds of 1e6-1e8: stay clear of restarts decided by rounding noise)


def make_inputs(K, N, seed, scale=None):
    """Shared with tests/test_vi_independent.py."""
    rng = np.random.default_rng(seed)
    if scale is not None:
        logl = np.full((K, N), np.log(0.01))
        for _ in range(3):
            logl[rng.integers(0, K, size=N), np.arange(N)] = -rng.random(N) * 3
        return logl, np.log(rng.integers(1, 40 * scale, size=N).astype(np.float64)), np.ones(K)
    logl = rng.normal(-6.0, 2.0, size=(K, N))
    logl[rng.integers(0, K, size=N), np.arange(N)] = rng.normal(-0.4, 0.2, size=N)     # every class has a likely group
    logl[rng.integers(0, K, size=N), np.arange(N)] = rng.normal(-1.0, 0.5, size=N)     # ... and often a competitor
    counts = rng.integers(1, 40, size=N).astype(np.float64)
    with np.errstate(divide="ignore"):
        lc = np.log(counts)
        if seed % 2 == 1:
            lc[::7] = -np.inf                                                           # bootstrap-style unobserved classes
    alpha0 = rng.uniform(0.5, 2.0, size=K) if seed % 3 == 0 else np.ones(K)
    return logl, lc, alpha0


def main():
    out = {}
    for name, spec in CASES.items():
        logl, lc, alpha0 = make_inputs(*spec)
        for algo in ("rcg", "em"):
            r = ref.run(algo, logl, lc, alpha0=alpha0, tol=TOL_BY_CASE.get(name, TOL), max_iters=400)
            for key in ("theta", "N_k", "trace_bound", "trace_gnorm", "trace_reset"):
                out[f"{name}_{algo}_{key}"] = r[key]
            out[f"{name}_{algo}_iters"] = np.array(r["iters"])
            out[f"{name}_{algo}_converged"] = np.array(r["converged"])
            print(name, algo, r["iters"], r["converged"], int(r["trace_reset"].sum()), r["bound"])
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "vi_independent.npz"), **out)


if __name__ == "__main__":
    main()
