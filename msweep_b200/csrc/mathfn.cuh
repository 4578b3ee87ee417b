// mathfn.cuh — small math kernels shared by host tests and device code.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define MSWB_HD __host__ __device__ __forceinline__
#else
#define MSWB_HD inline
#endif

namespace mswb {

#ifdef __CUDACC__
// Coefficients of exp_nonpos for the device: operands straight from the constant bank.  As immediates they are
// rematerialised with uniform moves around every use once registers are tight — a quarter of all issue slots of the
// RCG sweeps at K = 50 (ncu, profiles/small_k_r02.txt).
static __constant__ double EXPC[16] = {
    2.0876756987868098979e-09, 1.6059043836821614599e-10,   // 1/12!, 1/13!
    2.7557319223985890653e-07, 2.5052108385441718775e-08,   // 1/10!, 1/11!
    2.4801587301587301566e-05, 2.7557319223985892511e-06,   // 1/8!,  1/9!
    1.3888888888888889419e-03, 1.9841269841269841253e-04,   // 1/6!,  1/7!
    4.1666666666666664354e-02, 8.3333333333333332177e-03,   // 1/4!,  1/5!
    0.5, 1.6666666666666665741e-01,                         // 1/2!,  1/3!
    1.4426950408889634074, 6755399441055744.0,              // log2(e), 1.5 * 2^52
    -6.93147180369123816490e-01, -1.90821492927058770002e-10};   // -ln2 (two-term Cody-Waite)
#endif

MSWB_HD double bits_to_double(long long b) {
#ifdef __CUDA_ARCH__
  return __longlong_as_double(b);
#else
  double d; memcpy(&d, &b, sizeof(d)); return d;
#endif
}
MSWB_HD long long double_to_bits(double d) {
#ifdef __CUDA_ARCH__
  return __double_as_longlong(d);
#else
  long long b; memcpy(&b, &d, sizeof(b)); return b;
#endif
}

// exp(x) for x <= 0 (log-probabilities and max-shifted logits are never positive on this path).
// Branch-free core: n = round(x log2 e), r = x - n ln2 (two-term Cody-Waite), degree-13 Taylor in r
// (|r| <= 0.3466: truncation error 4e-18), result scaled by adding n to the exponent field.
// Below -707 (e^-707 = 9e-308, the edge of the normal range) the result is flushed to 0; -inf gives 0.
// Measured against libm over 4e6 arguments in [-707, 0]: max relative error 3.1e-16 (tests/test_mathfn.py).
MSWB_HD double exp_nonpos(double x) {
  const double xs = fmax(x, -707.0);                         // keeps the arithmetic finite; selected away below
#ifdef __CUDA_ARCH__
  {   // the same arithmetic as below, coefficients from the constant bank
    const double tn = fma(xs, EXPC[12], EXPC[13]);
    const double n = tn - EXPC[13];
    double r = fma(n, EXPC[14], xs);
    r = fma(n, EXPC[15], r);
    const double r2 = r * r;
    double pe = EXPC[0], po = EXPC[1];
    pe = fma(pe, r2, EXPC[2]); po = fma(po, r2, EXPC[3]);
    pe = fma(pe, r2, EXPC[4]); po = fma(po, r2, EXPC[5]);
    pe = fma(pe, r2, EXPC[6]); po = fma(po, r2, EXPC[7]);
    pe = fma(pe, r2, EXPC[8]); po = fma(po, r2, EXPC[9]);
    pe = fma(pe, r2, EXPC[10]); po = fma(po, r2, EXPC[11]);
    pe = fma(pe, r2, 1.0); po = fma(po, r2, 1.0);
    const double p = fma(po, r, pe);
    const double scale = __hiloint2double((__double2loint(tn) + 1023) << 20, 0);
    return x > -707.0 ? p * scale : 0.0;
  }
#endif
  const double magic = 6755399441055744.0;                   // 1.5 * 2^52: adding it rounds to nearest integer
  const double tn = fma(xs, 1.4426950408889634074, magic);
  const double n = tn - magic;
  double r = fma(n, -6.93147180369123816490e-01, xs);
  r = fma(n, -1.90821492927058770002e-10, r);
  // degree-13 Taylor, split in even/odd halves of r^2 to halve the dependent chain
  const double r2 = r * r;
  double pe = 2.0876756987868098979e-09;                     // 1/12!
  double po = 1.6059043836821614599e-10;                     // 1/13!
  pe = fma(pe, r2, 2.7557319223985890653e-07);               // 1/10!
  po = fma(po, r2, 2.5052108385441718775e-08);               // 1/11!
  pe = fma(pe, r2, 2.4801587301587301566e-05);               // 1/8!
  po = fma(po, r2, 2.7557319223985892511e-06);               // 1/9!
  pe = fma(pe, r2, 1.3888888888888889419e-03);               // 1/6!
  po = fma(po, r2, 1.9841269841269841253e-04);               // 1/7!
  pe = fma(pe, r2, 4.1666666666666664354e-02);               // 1/4!
  po = fma(po, r2, 8.3333333333333332177e-03);               // 1/5!
  pe = fma(pe, r2, 0.5);                                     // 1/2!
  po = fma(po, r2, 1.6666666666666665741e-01);               // 1/3!
  pe = fma(pe, r2, 1.0);
  po = fma(po, r2, 1.0);
  const double p = fma(po, r, pe);
  // the low 32 bits of tn hold n as a two's complement integer (magic-number rounding): 2^n built in the
  // high word only (n >= -1021 here, so the biased exponent stays positive)
#ifdef __CUDA_ARCH__
  const double scale = __hiloint2double((__double2loint(tn) + 1023) << 20, 0);
#else
  const long long ni = (long long)(int)(unsigned)(double_to_bits(tn) & 0xffffffffLL);
  const double scale = bits_to_double((ni + 1023) << 52);
#endif
  return x > -707.0 ? p * scale : 0.0;
}

} // namespace mswb
