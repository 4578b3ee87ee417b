"""CPU: the C-ABI library loads, exports every symbol include/msweep_b200.h declares, and refuses to
compute without a GPU (no fallback)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest


def test_library_exports_every_declared_symbol(mswb):
    names = mswb.declared_symbols()
    assert len(names) >= 25
    lib = mswb.lib()
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_no_undeclared_exports(mswb):
    out = subprocess.run(["nm", "-D", "--defined-only", mswb.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert exported == set(mswb.declared_symbols())


def test_header_is_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "msweep_b200.h"\nint main(void){ mswb_vi_opts o; (void)o; return 0; }\n')
    inc = os.path.join(os.path.dirname(__file__), "..", "include")
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", inc, "-c", str(src), "-o", str(tmp_path / "t.o")])


def test_version_and_error_strings(mswb):
    assert b"sm_100a" in mswb.lib().mswb_version()
    assert mswb.lib().mswb_last_error() is not None


def test_no_cpu_fallback(mswb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(mswb.MswbError, match="no CUDA device"):
        mswb.Context(0)


def test_struct_layouts_match_header(mswb):
    # ViOpts / ViStat mirror the C structs: sizes as the C compiler sees them
    import tempfile
    inc = os.path.join(os.path.dirname(__file__), "..", "include")
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "s.c")
        with open(src, "w") as f:
            f.write('#include <stdio.h>\n#include "msweep_b200.h"\nint main(void){printf("%zu %zu\\n", sizeof(mswb_vi_opts), sizeof(mswb_vi_stat));return 0;}\n')
        exe = os.path.join(d, "s")
        subprocess.check_call(["/usr/bin/gcc", "-I", inc, src, "-o", exe])
        a, b = map(int, subprocess.check_output([exe]).split())
    assert C.sizeof(mswb.ViOpts) == a and C.sizeof(mswb.ViStat) == b


def test_null_arguments_are_errors_not_crashes(mswb):
    """Every entry validates its pointers before touching the device: a NULL handle is an error code + message."""
    L = mswb.lib()
    out = C.c_void_p()
    a, b = C.c_uint64(), C.c_uint64()
    calls = [
        lambda: L.mswb_shard_range(None, C.c_uint64(10), C.byref(a), C.byref(b)),
        lambda: L.mswb_ec_build(None, C.c_uint64(0), C.c_uint64(1), None, None, C.byref(out)),
        lambda: L.mswb_ec_info(None, None, None, None, None),
        lambda: L.mswb_lik_build(None, None, None, C.c_uint32(1), None, C.c_double(0.65), C.c_double(0.01), C.c_double(0.01), C.c_uint64(0), C.c_int(0), C.byref(out)),
        lambda: L.mswb_lik_info(None, None, None, None, None, None),
        lambda: L.mswb_lik_mask(None, None, None),
        lambda: L.mswb_vi_begin(None, None, None, None, None, C.byref(out)),
        lambda: L.mswb_vi_step(None, C.c_uint64(1)),
        lambda: L.mswb_vi_poll(None, None),
        lambda: L.mswb_vi_finish(None, None, None, None),
        lambda: L.mswb_vi_posteriors(None, None, C.c_uint64(0), C.c_uint64(0), None),
        lambda: L.mswb_bootstrap_resample(None, None, C.c_int32(1), C.c_uint64(0), C.c_int(0), C.c_uint64(1), None),
        lambda: L.mswb_ctx_sync(None),
    ]
    for call in calls:
        assert call() != 0
        assert len(L.mswb_last_error()) > 0
    L.mswb_ctx_destroy(None); L.mswb_aln_destroy(None); L.mswb_lik_destroy(None)      # destroying nothing is fine


def test_pattern_hash_host_entry(mswb, oracle):
    rng = np.random.default_rng(5)
    assert mswb.pattern_hash([0]) == 0x517cc1b727220a95 and mswb.pattern_hash([]) == 0
    for _ in range(50):
        t = np.sort(rng.choice(60000, size=int(rng.integers(1, 90)), replace=False))
        assert mswb.pattern_hash(t) == oracle.pattern_hash(t)
