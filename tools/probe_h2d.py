"""EC build from PAGEABLE host CSR buffers at config 2's size (1e7 reads): staged multi-threaded H2D (default) against the plain
cudaMemcpyAsync (MSWB_H2D_STAGED=0, read once per process).  Prints seconds per build and the CSR size."""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import msweep_b200 as M
from msweep_b200 import synth

wl = synth.generate_ec_patterns(1_000_000, 1000, 60, n_present=20, seed=55, dup_factor=9.0)
ctx = M.Context(0)
out = {"staged": os.environ.get("MSWB_H2D_STAGED", "1"), "csr_gb": round((wl.row_ptr.nbytes + wl.targets.nbytes) / 1e9, 3), "reads": int(wl.n_reads)}
ts = []
for _ in range(4):
    ctx.sync(); t0 = time.perf_counter()
    aln = M.Alignment(ctx, wl.n_reads, wl.n_targets, wl.row_ptr, wl.targets)
    ctx.sync(); ts.append(round(time.perf_counter() - t0, 4))
    n = aln.n_ecs
    aln.close()
out["ec_build_s"] = ts
out["ecs"] = int(n)
print(json.dumps(out))
