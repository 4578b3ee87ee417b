"""Regenerates tests/golden/bench_check.npz: the oracle's answers on the seeded case bench.py re-runs before every
timed region (`check.parity` at 1 GPU, `check.multi_gpu_parity` at N > 1, EC-sharded and hash-partitioned).
bench.py's own arm never executes anything under oracle/: it compares the CUDA results with these frozen numbers.

Run from the repo root:  python tests/golden/make_bench_check.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from msweep_b200 import synth            # noqa: E402
from oracle import pyoracle as orc       # noqa: E402

CASE = dict(n_reads=60000, n_targets=3000, n_groups=50, n_present=5, n_templates=400, p_noise=0.02, seed=5)
MIN_HITS = 50


def main():
    wl = synth.generate(**CASE)
    ec = orc.ec_build_csr(wl.n_reads, wl.n_targets, wl.row_ptr, wl.targets)
    lik = orc.lik_build(ec, wl.group_of_target, wl.group_sizes)
    lik_mh = orc.lik_build(ec, wl.group_of_target, wl.group_sizes, min_hits=MIN_HITS)
    out = {"case": np.array(repr(CASE)), "min_hits": np.array(MIN_HITS), "n_ecs": np.array(ec.n_ecs),
           "n_aligned": np.array(int(ec.count.sum())),
           "hash_xor": np.bitwise_xor.reduce(ec.hash), "hash_sum": np.array(int(ec.hash.astype(object).sum()) % (1 << 64), np.uint64),
           "count_dot": np.array(int((ec.count.astype(object) * np.arange(1, ec.n_ecs + 1).astype(object)).sum()) % (1 << 64), np.uint64),
           "mask_mh": lik_mh.mask, "hits_mh": lik_mh.hits}
    for name, algo, L in (("rcg", "rcg", lik), ("em", "em", lik), ("rcg_mh", "rcg", lik_mh)):
        r = orc.vi_run(algo, L.logl, L.log_counts)
        out[name + "_theta"] = r.theta
        out[name + "_bound"] = np.array(r.bound)
        out[name + "_iters"] = np.array(r.iters)
        out[name + "_resets"] = np.array(int(r.trace_reset.sum()))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "bench_check.npz"), **out)
    print({k: (v.shape if hasattr(v, "shape") and v.shape else v) for k, v in out.items() if not k.endswith("theta")})


if __name__ == "__main__":
    main()
