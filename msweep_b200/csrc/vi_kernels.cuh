// vi_kernels.cuh — the fused VI sweeps (K6 / K7 of SURVEY.md §2.4), EC-major, HBM-bound.
//
// Every sweep has the same shape.  A CTA of NT threads is split into G = NT/TPR row groups; a group of
// TPR threads owns R consecutive classes (rows) per batch and strides the K columns of a row with
// 128-bit loads: thread t of the group holds columns VEC*(t + TPR*i) .. +VEC-1, i < KITER.  Per-class
// quantities (the logsumexp / the normaliser S_j / the mean step) are reductions ALONG a row:
// warp shuffles, plus one shared-memory hop when a row spans several warps.  Per-group quantities
// (the expected counts N_k) accumulate DOWN the rows in registers and leave the CTA once, as a
// per-CTA partial vector that a second tiny kernel sums in a fixed order (no atomics: results are
// bit-reproducible and identical however many CTAs ran).  The grid is persistent (a multiple of the
// SM count) and walks the row batches with a grid stride.
#pragma once
#include "common.cuh"

namespace mswb {

// Device-resident control block of one optimisation: the host never has to be in the loop.
struct ViCtl {
  double bound, oldbound;
  double oldnorm, newnorm, beta;
  double bound_const, tol, sum_counts, dg_max;
  unsigned long long iter, max_iters, resets;
  int use_old, didreset, converged, done, fault;
};

template <int TPR_, int KITER_, int R_> struct Tile {
  static constexpr int TPR = TPR_, KITER = KITER_, R = R_;
  static constexpr int NT = TPR < 256 ? 256 : TPR;
  static constexpr int G = NT / TPR;
  static constexpr int NW = NT / 32;
  static constexpr int WPG = TPR / 32;   // warps per row group
};

// ---- reductions along a row -------------------------------------------------------------------------
// scratch: 2 * NW * R doubles.  Two buffers alternate so that one __syncthreads per call is enough.
template <class TL, bool IS_MAX> __device__ __forceinline__ void group_reduce(double (&v)[TL::R], double *scratch, int &phase) {
#pragma unroll
  for (int r = 0; r < TL::R; ++r) v[r] = IS_MAX ? warp_max(v[r]) : warp_sum(v[r]);
  if constexpr (TL::WPG > 1) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *buf = scratch + phase * (TL::NW * TL::R);
    phase ^= 1;
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < TL::R; ++r) buf[warp * TL::R + r] = v[r];
    }
    __syncthreads();
    const int w0 = (warp / TL::WPG) * TL::WPG;
#pragma unroll
    for (int r = 0; r < TL::R; ++r) {
      double a = buf[w0 * TL::R + r];
#pragma unroll
      for (int w = 1; w < TL::WPG; ++w) a = IS_MAX ? fmax(a, buf[(w0 + w) * TL::R + r]) : a + buf[(w0 + w) * TL::R + r];
      v[r] = a;
    }
  }
}

// ---- per-CTA partial vector: combine the G row groups, then one coalesced store ---------------------
// acc[i][v] belongs to column VEC*(t + TPR*i) + v.  comb: G * TPR*KITER*VEC doubles when G > 1.
template <class TL, int VEC>
__device__ __forceinline__ void store_partials(const double (&acc)[TL::KITER][VEC], double *comb, double *out, int K) {
  const int t = threadIdx.x % TL::TPR, g = threadIdx.x / TL::TPR;
  constexpr int KCAP = TL::TPR * TL::KITER * VEC;
  if constexpr (TL::G == 1) {
#pragma unroll
    for (int i = 0; i < TL::KITER; ++i)
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const int k = VEC * (t + TL::TPR * i) + v;
        if (k < K) out[k] = acc[i][v];
      }
  } else {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < TL::KITER; ++i)
#pragma unroll
      for (int v = 0; v < VEC; ++v) comb[g * KCAP + VEC * (t + TL::TPR * i) + v] = acc[i][v];
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += TL::NT) {
      double a = 0.0;
#pragma unroll
      for (int gg = 0; gg < TL::G; ++gg) a += comb[gg * KCAP + k];
      out[k] = a;
    }
  }
}

// =====================================================================================================
// EM / VB pass, linear domain.  P(j,k) = exp(logl(j,k) - M_j) is stored once; a pass is two GEMVs that
// share one read of P:   S_j = sum_k P(j,k) w_k ,  A_k = sum_j P(j,k) c_j / S_j ,  with
// w_k = exp(digamma(N_k) - max digamma).  Then N_k = alpha0_k + w_k A_k and the data term of the ELBO
// is sum_j c_j (log S_j + M_j) (+ constants applied by the control kernel).  Two FMAs per element.
// =====================================================================================================
template <typename ST> struct VecOf;
template <> struct VecOf<double> { using type = double2; static constexpr int VEC = 2; };
template <> struct VecOf<float> { using type = float4; static constexpr int VEC = 4; };

__device__ __forceinline__ void unpack(const double2 &x, double (&o)[2]) { o[0] = x.x; o[1] = x.y; }
__device__ __forceinline__ void unpack(const float4 &x, float (&o)[4]) { o[0] = x.x; o[1] = x.y; o[2] = x.z; o[3] = x.w; }

template <typename ST, class TL>
__global__ void __launch_bounds__(TL::NT)
em_lin_pass_kernel(const ST *__restrict__ P, int ld, const double *__restrict__ rowmax, const double *__restrict__ counts,
                   const double *__restrict__ w, const ViCtl *__restrict__ ctl, double *__restrict__ partials,
                   int pstride, unsigned long long N, int K) {
  using VT = typename VecOf<ST>::type;
  constexpr int VEC = VecOf<ST>::VEC;
  constexpr int R = TL::R, KITER = TL::KITER, TPR = TL::TPR;
  if (ctl->done) return;
  __shared__ double s_red[2 * TL::NW * R];
  __shared__ double s_comb[TL::G > 1 ? TL::G * TPR * KITER * VEC : 1];
  __shared__ double s_blk[32];
  const int t = threadIdx.x % TPR, g = threadIdx.x / TPR;

  ST wv[KITER][VEC];
#pragma unroll
  for (int i = 0; i < KITER; ++i)
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const int k = VEC * (t + TPR * i) + v;
      wv[i][v] = k < K ? (ST)w[k] : (ST)0;
    }
  double acc[KITER][VEC];
#pragma unroll
  for (int i = 0; i < KITER; ++i)
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[i][v] = 0.0;
  double elbo = 0.0;
  int fault = 0;
  int phase = 0;
  const unsigned long long rows_per_batch = (unsigned long long)TL::G * R;
  const unsigned long long n_batches = (N + rows_per_batch - 1) / rows_per_batch;
  const int nvec = ld / VEC;

  for (unsigned long long b = blockIdx.x; b < n_batches; b += gridDim.x) {
    const unsigned long long row0 = b * rows_per_batch + (unsigned long long)g * R;
    ST pv[R][KITER][VEC];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const unsigned long long row = row0 + r;
      const VT *rp = reinterpret_cast<const VT *>(P + row * (unsigned long long)ld);
#pragma unroll
      for (int i = 0; i < KITER; ++i) {
        const int idx = t + TPR * i;
        if (row < N && idx < nvec) unpack(ld_stream(rp + idx), pv[r][i]);
        else {
#pragma unroll
          for (int v = 0; v < VEC; ++v) pv[r][i][v] = (ST)0;
        }
      }
    }
    double s[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      ST a = (ST)0;   // fp32 storage: the row-local dot product runs in fp32, everything across rows in fp64
#pragma unroll
      for (int i = 0; i < KITER; ++i)
#pragma unroll
        for (int v = 0; v < VEC; ++v) a = fma(pv[r][i][v], wv[i][v], a);
      s[r] = (double)a;
    }
    group_reduce<TL, false>(s, s_red, phase);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const unsigned long long row = row0 + r;
      if (row < N) {
        const double c = counts[row];
        if (c > 0.0) {
          if (!(s[r] > 0.0) || isinf(s[r])) { fault = 1; continue; }
          const double inv = c / s[r];
#pragma unroll
          for (int i = 0; i < KITER; ++i)
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[i][v] = fma((double)pv[r][i][v], inv, acc[i][v]);
          if (t == 0) elbo += c * (log(s[r]) + rowmax[row]);
        }
      }
    }
  }
  double *out = partials + (unsigned long long)blockIdx.x * pstride;
  store_partials<TL, VEC>(acc, s_comb, out, K);
  elbo = block_sum<TL::NT>(elbo, s_blk);
  if (threadIdx.x == 0) { out[K] = elbo; out[K + 1] = 0.0; }
  if (fault) atomicExch(const_cast<int *>(&ctl->fault), 1);
}

// =====================================================================================================
// Log-domain sweeps (fp64): the RCG optimiser, the restart step, and the EM fallback.
// =====================================================================================================

// Sweep A of an RCG iteration ("mixt_negnatgrad"): d = logl + (digamma(N_k) - 1) - gamma,
// newnorm = sum_jk q (d - <d>_j) d  with q = exp(gamma), <d>_j = sum_k q d.  Nothing is written: d is
// recomputed by sweep B, which saves 16 B/element of traffic over storing it.
template <class TL>
__global__ void __launch_bounds__(TL::NT)
rcg_sweep_a_kernel(const double *__restrict__ logl, const double *__restrict__ gamma, int ld,
                   const double *__restrict__ dgm1, const ViCtl *__restrict__ ctl, double *__restrict__ partials,
                   int pstride, unsigned long long N, int K) {
  constexpr int R = TL::R, KITER = TL::KITER, TPR = TL::TPR;
  if (ctl->done) return;
  __shared__ double s_red[2 * TL::NW * R];
  __shared__ double s_blk[32];
  const int t = threadIdx.x % TPR, g = threadIdx.x / TPR;
  double dk[KITER][2];
  bool ok[KITER][2];
#pragma unroll
  for (int i = 0; i < KITER; ++i)
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      const int k = 2 * (t + TPR * i) + v;
      ok[i][v] = k < K;
      dk[i][v] = ok[i][v] ? dgm1[k] : 0.0;
    }
  double nn = 0.0;
  int phase = 0;
  const unsigned long long rows_per_batch = (unsigned long long)TL::G * R;
  const unsigned long long n_batches = (N + rows_per_batch - 1) / rows_per_batch;
  const int nvec = ld / 2;
  for (unsigned long long b = blockIdx.x; b < n_batches; b += gridDim.x) {
    const unsigned long long row0 = b * rows_per_batch + (unsigned long long)g * R;
    double d[R][KITER][2], q[R][KITER][2];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const unsigned long long row = row0 + r;
      const double2 *lp = reinterpret_cast<const double2 *>(logl + row * (unsigned long long)ld);
      const double2 *gp = reinterpret_cast<const double2 *>(gamma + row * (unsigned long long)ld);
#pragma unroll
      for (int i = 0; i < KITER; ++i) {
        const int idx = t + TPR * i;
        if (row < N && idx < nvec) { unpack(ld_stream(lp + idx), d[r][i]); unpack(ld_stream(gp + idx), q[r][i]); }
        else { d[r][i][0] = d[r][i][1] = 0.0; q[r][i][0] = q[r][i][1] = 0.0; }
      }
    }
    double s[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const bool rv = row0 + r < N;
      s[r] = 0.0;
#pragma unroll
      for (int i = 0; i < KITER; ++i)
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          const bool on = rv && ok[i][v];
          const double gam = q[r][i][v];
          const double dd = on ? d[r][i][v] + dk[i][v] - gam : 0.0;
          const double qq = on ? exp(gam) : 0.0;
          d[r][i][v] = dd; q[r][i][v] = qq;
          s[r] = fma(dd, qq, s[r]);
        }
    }
    group_reduce<TL, false>(s, s_red, phase);
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int i = 0; i < KITER; ++i)
#pragma unroll
        for (int v = 0; v < 2; ++v) nn = fma(q[r][i][v] * (d[r][i][v] - s[r]), d[r][i][v], nn);
  }
  nn = block_sum<TL::NT>(nn, s_blk);
  if (threadIdx.x == 0) partials[(unsigned long long)blockIdx.x * pstride + K + 1] = nn;
}

// Sweep B of an RCG iteration: step = d (+ beta * oldstep), gamma += step, renormalise every class,
// store gamma and step, accumulate N_k - alpha0 = sum_j c_j q and the data term of the ELBO.
// MODE 0: RCG step.  MODE 1: plain step from the current digamma vector, gamma = normalise(logl + dg)
// (the RCG restart, and the log-domain EM pass); WRITE says whether gamma is stored.
template <class TL, int MODE, bool WRITE>
__global__ void __launch_bounds__(TL::NT)
rcg_sweep_b_kernel(const double *__restrict__ logl, double *__restrict__ gamma, double *__restrict__ step, int ld,
                   const double *__restrict__ dgv, const double *__restrict__ counts, const ViCtl *__restrict__ ctl,
                   double *__restrict__ partials, int pstride, unsigned long long N, int K, int only_if_reset) {
  constexpr int R = TL::R, KITER = TL::KITER, TPR = TL::TPR;
  if (ctl->done) return;
  if (only_if_reset && !ctl->didreset) return;
  __shared__ double s_red[2 * TL::NW * R];
  __shared__ double s_comb[TL::G > 1 ? TL::G * TPR * KITER * 2 : 1];
  __shared__ double s_blk[32];
  const int t = threadIdx.x % TPR, g = threadIdx.x / TPR;
  const double beta = ctl->beta;
  const bool use_old = MODE == 0 && ctl->use_old != 0;
  double dk[KITER][2];
  bool ok[KITER][2];
#pragma unroll
  for (int i = 0; i < KITER; ++i)
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      const int k = 2 * (t + TPR * i) + v;
      ok[i][v] = k < K;
      dk[i][v] = ok[i][v] ? dgv[k] : 0.0;
    }
  double acc[KITER][2];
#pragma unroll
  for (int i = 0; i < KITER; ++i) acc[i][0] = acc[i][1] = 0.0;
  double bound = 0.0;
  int phase = 0;
  const unsigned long long rows_per_batch = (unsigned long long)TL::G * R;
  const unsigned long long n_batches = (N + rows_per_batch - 1) / rows_per_batch;
  const int nvec = ld / 2;
  const double NEG_INF = -INFINITY;

  for (unsigned long long b = blockIdx.x; b < n_batches; b += gridDim.x) {
    const unsigned long long row0 = b * rows_per_batch + (unsigned long long)g * R;
    double l[R][KITER][2], gn[R][KITER][2];
    {
      double st[R][KITER][2];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const unsigned long long row = row0 + r;
        const unsigned long long off = row * (unsigned long long)ld;
        const double2 *lp = reinterpret_cast<const double2 *>(logl + off);
        const double2 *gp = reinterpret_cast<const double2 *>(gamma + off);
        const double2 *sp = reinterpret_cast<const double2 *>(step + off);
#pragma unroll
        for (int i = 0; i < KITER; ++i) {
          const int idx = t + TPR * i;
          const bool in = row < N && idx < nvec;
          if (in) unpack(ld_stream(lp + idx), l[r][i]); else l[r][i][0] = l[r][i][1] = 0.0;
          if (MODE == 0) {
            if (in) unpack(ld_stream(gp + idx), gn[r][i]); else gn[r][i][0] = gn[r][i][1] = 0.0;
            if (in && use_old) unpack(ld_stream(sp + idx), st[r][i]); else st[r][i][0] = st[r][i][1] = 0.0;
          }
        }
      }
      // the new direction leaves for HBM straight away so that its registers die before the reductions
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const unsigned long long row = row0 + r;
        double2 *sp = reinterpret_cast<double2 *>(step + row * (unsigned long long)ld);
#pragma unroll
        for (int i = 0; i < KITER; ++i) {
          const int idx = t + TPR * i;
#pragma unroll
          for (int v = 0; v < 2; ++v) {
            const bool on = row < N && ok[i][v];
            double g2;
            if (MODE == 0) {
              const double d = l[r][i][v] + dk[i][v] - gn[r][i][v];
              const double s = use_old ? fma(beta, st[r][i][v], d) : d;
              st[r][i][v] = on ? s : 0.0;
              g2 = gn[r][i][v] + s;
            } else {
              g2 = l[r][i][v] + dk[i][v];
            }
            gn[r][i][v] = on ? g2 : NEG_INF;
          }
          if (MODE == 0 && row < N && idx < nvec) st_stream(sp + idx, make_double2(st[r][i][0], st[r][i][1]));
        }
      }
    }
    double m[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      m[r] = NEG_INF;
#pragma unroll
      for (int i = 0; i < KITER; ++i) m[r] = fmax(m[r], fmax(gn[r][i][0], gn[r][i][1]));
    }
    group_reduce<TL, true>(m, s_red, phase);
    double e[R][KITER][2], sum[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const double mm = m[r] == NEG_INF ? 0.0 : m[r];   // a row past the end: keep the arithmetic finite
      sum[r] = 0.0;
#pragma unroll
      for (int i = 0; i < KITER; ++i)
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          gn[r][i][v] -= mm;
          e[r][i][v] = exp(gn[r][i][v]);                 // exp(-inf) = 0 on padding columns
          sum[r] += e[r][i][v];
        }
    }
    group_reduce<TL, false>(sum, s_red, phase);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const unsigned long long row = row0 + r;
      if (row >= N) continue;
      const double lsum = log(sum[r]);
      const double inv = 1.0 / sum[r];
      const double c = counts[row];
      double2 *gp = reinterpret_cast<double2 *>(gamma + row * (unsigned long long)ld);
#pragma unroll
      for (int i = 0; i < KITER; ++i) {
        const int idx = t + TPR * i;
        if (idx >= nvec) continue;
        double gout[2];
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          const double g2 = gn[r][i][v] - lsum;          // normalised log-responsibility
          gout[v] = ok[i][v] ? g2 : 0.0;
          if (ok[i][v] && c > 0.0) {
            const double cq = c * (e[r][i][v] * inv);
            acc[i][v] += cq;
            if (cq > 0.0) bound = fma(cq, l[r][i][v] - g2, bound);
          }
        }
        if (WRITE) st_stream(gp + idx, make_double2(gout[0], gout[1]));
      }
    }
  }
  double *out = partials + (unsigned long long)blockIdx.x * pstride;
  store_partials<TL, 2>(acc, s_comb, out, K);
  bound = block_sum<TL::NT>(bound, s_blk);
  if (threadIdx.x == 0) out[K] = bound;
}

// ---- small kernels --------------------------------------------------------------------------------
// red[v] = sum over CTAs of partials[cta][v], fixed order.  v < nvals.
__global__ void finalize_partials_kernel(const double *__restrict__ partials, int pstride, int n_ctas, int nvals,
                                         double *__restrict__ red, const ViCtl *__restrict__ ctl, int only_if_reset) {
  if (ctl->done) return;
  if (only_if_reset && !ctl->didreset) return;
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nvals) return;
  double a = 0.0;
  for (int c = 0; c < n_ctas; ++c) a += partials[(size_t)c * pstride + v];
  red[v] = a;
}

} // namespace mswb
