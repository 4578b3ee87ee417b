"""The optimiser against an INDEPENDENT restatement (tests/ref_numpy_vi.py: numpy, class-major, long-double sums, the
restart derived algebraically) frozen in tests/golden/vi_independent.npz.  rcgpar itself is off-tree, so this does not pin
the reference's arithmetic — parity of the optimiser stays "unpinned"  — but it removes the single point of
failure: the oracle (oracle/oracle.cpp) and the CUDA path both have to follow a second, separately written
implementation of the same equations, iteration by iteration."""
import os

import numpy as np
import pytest

from tests.golden.make_vi_independent import CASES, TOL, TOL_BY_CASE, make_inputs

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "vi_independent.npz"))
THETA_TOL, ELBO_RTOL = 1e-6, 1e-9


def _check(name, algo, iters, converged, theta, tb, tg, tr):
    g = lambda key: GOLD[f"{name}_{algo}_{key}"]
    assert iters == int(g("iters")) and bool(converged) == bool(g("converged"))
    assert np.array_equal(np.asarray(tr, np.uint8), g("trace_reset"))                     # the same restart pattern
    assert np.max(np.abs(tb - g("trace_bound")) / np.abs(g("trace_bound"))) <= ELBO_RTOL  # the bound, iteration by iteration
    if algo == "rcg":
        # (the norm is a difference of nearly equal sums: near the optimum only its absolute size is meaningful)
        assert np.allclose(tg, g("trace_gnorm"), rtol=1e-6, atol=1e-12 * (1.0 + g("trace_gnorm")[0]))
    assert np.max(np.abs(theta - g("theta"))) < THETA_TOL


@pytest.mark.parametrize("algo", ["rcg", "em"])
@pytest.mark.parametrize("name", list(CASES))
def test_oracle_follows_the_independent_restatement(oracle, name, algo):
    if name.endswith("restart") and algo == "rcg":
        assert GOLD[f"{name}_rcg_trace_reset"][0] == 1, "these cases are here for the restart branch"
    logl, lc, alpha0 = make_inputs(*CASES[name])
    r = oracle.vi_run(algo, logl, lc, alpha0=alpha0, tol=TOL_BY_CASE.get(name, TOL), max_iters=400)
    _check(name, algo, r.iters, r.converged, r.theta, r.trace_bound, r.trace_gnorm, r.trace_reset)


@pytest.mark.parametrize("name", list(CASES))
def test_the_restatement_is_reproducible(name):
    """The fixture is what tests/ref_numpy_vi.py computes today (a stale fixture would hide a drift)."""
    from tests import ref_numpy_vi as ref
    logl, lc, alpha0 = make_inputs(*CASES[name])
    r = ref.run("rcg", logl, lc, alpha0=alpha0, tol=TOL_BY_CASE.get(name, TOL), max_iters=400)
    assert r["iters"] == int(GOLD[f"{name}_rcg_iters"])
    assert np.allclose(r["theta"], GOLD[f"{name}_rcg_theta"], rtol=0, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("algo", ["rcg", "em"])
@pytest.mark.parametrize("name", list(CASES))
def test_cuda_follows_the_independent_restatement(mswb, ctx, name, algo):
    logl, lc, alpha0 = make_inputs(*CASES[name])
    lik = mswb.Likelihood.from_dense(ctx, logl, np.where(np.isfinite(lc), lc, 0.0))
    sess = lik.vi_begin(mswb.ALGO_RCG if algo == "rcg" else mswb.ALGO_EM, alpha0=alpha0, log_counts=lc, tol=TOL_BY_CASE.get(name, TOL), max_iters=400)
    while True:
        sess.step(8)
        st = sess.poll()
        if st.converged or st.iters >= 400:
            break
    tb, tg, tr = sess.trace()
    res = sess.finish()
    _check(name, algo, res.iters, res.converged, res.theta, tb, tg, tr)
