"""Prints a digest of the EC table built from a pageable ~150 MB CSR (twice: cold and warm lanes).  Run once with
MSWB_H2D_STAGED_MIN_MB=1 (staged multi-threaded host-to-device copy) and once with MSWB_H2D_STAGED=0 (plain copy): the
digests must be equal (tests/test_gpu_scale.py)."""
import hashlib, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import msweep_b200 as M
from msweep_b200 import synth

wl = synth.generate_ec_patterns(400_000, 500, 40, n_present=30, seed=91, dup_factor=3.0)
assert wl.targets.nbytes >= 64 << 20, wl.targets.nbytes
ctx = M.Context(0)
h = hashlib.sha256()
for _ in range(2):
    aln = M.Alignment(ctx, wl.n_reads, wl.n_targets, wl.row_ptr, wl.targets)
    e = aln.export()
    for a in (e.hash, e.count, e.rep_read, e.pat_ptr, e.pat_targets, e.read_ptr, e.read_ids):
        h.update(np.ascontiguousarray(a).tobytes())
    aln.close()
print("digest", h.hexdigest(), "csr_mb", (wl.targets.nbytes + wl.row_ptr.nbytes) >> 20)
