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
