"""The synthetic workload generators (msweep_b200/synth.py) produce what the configs in BASELINE.json describe."""
import numpy as np

from msweep_b200 import synth


def test_ec_pattern_generator_shapes_and_order():
    wl = synth.generate_ec_patterns(5000, 40, 6, n_present=5, seed=3)
    assert wl.n_reads == 5000 and wl.n_targets == 240 and wl.n_groups == 40
    assert wl.row_ptr.dtype == np.uint64 and wl.targets.dtype == np.uint32 and wl.row_ptr[-1] == wl.targets.size
    for r in range(0, 5000, 97):                                   # rows ascending and unique: the C ABI's input contract
        row = wl.targets[int(wl.row_ptr[r]):int(wl.row_ptr[r + 1])]
        assert np.all(np.diff(row.astype(np.int64)) > 0)
    assert abs(wl.truth.sum() - 1.0) < 1e-12 and np.count_nonzero(wl.truth) == 5


def test_lineage_pool_leaves_the_other_lineages_empty():
    """config 4: most lineages receive no hit at all, so that --min-hits 1 prunes them."""
    wl = synth.generate_ec_patterns(20000, 500, 4, n_present=10, n_pool=25, seed=4)
    hit = np.unique(wl.group_of_target[np.unique(wl.targets)])
    assert hit.size <= 25
    assert set(np.flatnonzero(wl.truth)) <= set(hit.tolist())
    free = synth.generate_ec_patterns(20000, 500, 4, n_present=10, seed=4)
    assert np.unique(free.group_of_target[np.unique(free.targets)]).size > 400
