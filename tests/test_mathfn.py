"""exp_nonpos (msweep_b200/csrc/mathfn.cuh) against libm's long-double exp: the header is plain C++ on the
host, so the arithmetic the kernels run (same fused multiply-adds, same table) is checked here without a GPU."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include "mathfn.cuh"
#include <cstdio>
#include <cmath>
#include <random>
int main() {
  std::mt19937_64 g(1);
  std::uniform_real_distribution<double> wide(-707.0, 0.0), near(-40.0, 0.0);
  double w1 = 0;
  for (int i = 0; i < 4000000; ++i) {
    const double x = (i & 1) ? wide(g) : near(g);
    const long double ref = expl((long double)x);
    w1 = fmax(w1, (double)fabsl((mswb::exp_nonpos(x) - ref) / ref));
  }
  std::printf("%.6e\n", w1);
  std::printf("%.17g %.17g %.17g %.17g\n", mswb::exp_nonpos(0.0), mswb::exp_nonpos(-INFINITY), mswb::exp_nonpos(-707.0), mswb::exp_nonpos(-708.5));
  return 0;
}
'''


def test_exp_kernels_against_libm(tmp_path):
    src = tmp_path / "mathfn_check.cpp"
    src.write_text(SRC)
    exe = tmp_path / "mathfn_check"
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "msweep_b200", "csrc"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines()
    assert float(out[0]) < 3.5e-16          # ~1.5 ulp
    for line in out[1:]:
        one, zero_inf, edge, below = (float(x) for x in line.split())
        assert one == 1.0 and zero_inf == 0.0 and edge == 0.0 and below == 0.0     # flushed to 0 at and below -707
