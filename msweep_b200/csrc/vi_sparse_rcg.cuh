// vi_sparse_rcg.cuh — the RCG optimiser (rcgpar::rcg_optl_*) on the SPARSE likelihood storage.
//
// LL_WOR21 gives every group a class does NOT hit the same value l0 = log(zero_inflation) (include/Likelihood.hpp:96,
// 176-185).  RCG starts from the uniform gamma = log(1/K), and every update it applies to an entry (j, k) is
//     d_jk = logl_jk + (digamma(N_k) - 1) - gamma_jk,   step = d + beta * oldstep,   gamma += step - m_j.
// On the non-hit entries logl is the constant l0, so by induction gamma and the search direction stay SEPARABLE there:
//     gamma_jk = a_k + b_j ,   step_jk = u_k + v_j        for every group k that class j does not hit,
// exactly (not an approximation): d_jk = (psi_k - a_k) + (l0 - b_j), and sums of separable terms are separable.
// The dense K x N state of RCG — 32 bytes per (class, group) pair, the reason RCG could not run on config 3 at all —
// collapses to two K-vectors (a, u), two N-vectors (b, v) and explicit (gamma, step) values for the hits only.  Every
// row reduction the sweeps need splits into a closed form over the non-hit groups, built from global moments of the
// group vectors minus the class's few hit groups, plus an explicit sum over its hits:
//     sum_k exp(gamma_jk)             = exp(b_j) (M0 - sum_H E_k) + sum_hits exp(g_e),           E_k = exp(a_k), M0 = sum_k E_k
//     sum_k q (d - <d>)  d            = S2_j - S1_j^2  with the direction centred per class (no cancellation at the optimum)
//     N_k - alpha0_k                  = E_k (B - sum_{j hits k} c_j exp(b_j)) + sum_{j hits k} c_j q_jk,  B = sum_j c_j exp(b_j)
//     sum_k c_j q (logl - gamma)      = c_j exp(b_j) ((l0 - b_j)(M0 - sum_H E) - (Ma - sum_H E a)) + hits,  Ma = sum_k E_k a_k
// The iteration is the dense one, step for step (same Fletcher-Reeves ratio, same restart rule, same stopping rule): the
// trajectory differs from the dense sweeps only by summation order.  Bytes per iteration: ~150 per class + ~50 per hit,
// against 56 K per class.
//
// Kernels: rcgs_sweep_a_kernel (gradient norm), rcgs_sweep_b_kernel<MODE> (the K-sized group part of the step in its
// prologue, then step / restart, renormalise, N_k, bound), and the control step.  One warp per chunk of 32 classes, hits read hit-parallel (coalesced) and parked in shared memory, class sums
// class-parallel in list order, per-group sums in the fixed-point accumulators of the sparse EM pass (bit-reproducible).
#pragma once
#include "vi_kernels.cuh"

namespace mswb {

// Per-group vectors and global moments of sparse RCG (device pointers; identical on every rank).
struct RcgsGroup {
  double *a, *u;          // [K] group part of gamma (max-normalised: max_k a_k = 0) and of the search direction
  double *a_new, *u_new;  // [K] the same for the step being taken; committed by the control step when it is accepted
  double *ec;             // [K] e_k - ebar with e_k = psi_k - a_k   (sweep A)
  double *mom;            // [RCGS_MOM] see below
};
// mom[]: moments of the committed state (sweep A)
constexpr int RM_M0 = 0, RM_EBAR = 1, RM_M2C = 2, RCGS_MOM = 4;
// slots of the reduced vector behind the K per-group sums (the norm sits last so that the all-reduce after sweep B,
// K + 2 values, leaves it alone)
constexpr int RS_BOUND = 0, RS_MASS = 1, RS_NORM = 2;

// Moments for sweep A from the committed (a, psi): E_k = exp(a_k), M0, ebar = sum E e / M0, ec = e - ebar, M2c = sum E ec^2.
template <int NT>
__device__ __forceinline__ void rcgs_prep_a(const ViArrays &va, const RcgsGroup &g, int K, double *scratch) {
  double m0 = 0.0, m1 = 0.0;
  for (int k = threadIdx.x; k < K; k += NT) {
    const double E = exp(g.a[k]), e = va.dg[k] - g.a[k];
    m0 += E; m1 = fma(E, e, m1);
  }
  m0 = block_sum<NT>(m0, scratch);
  m1 = block_sum<NT>(m1, scratch);
  const double ebar = m1 / m0;
  double m2 = 0.0;
  for (int k = threadIdx.x; k < K; k += NT) {
    const double E = exp(g.a[k]), ec = va.dg[k] - g.a[k] - ebar;
    g.ec[k] = ec;
    m2 = fma(E * ec, ec, m2);
  }
  m2 = block_sum<NT>(m2, scratch);
  if (threadIdx.x == 0) { g.mom[RM_M0] = m0; g.mom[RM_EBAR] = ebar; g.mom[RM_M2C] = m2; }
}

// The group part of the step about to be taken, computed by EVERY CTA of sweep B in its prologue (K-sized: a few
// microseconds; one launch less per iteration than a kernel of its own) into shared memory; CTA 0 also publishes it for
// the control step.  MODE 0: u' = e + beta u, a' = a + u'.  MODE 1 (restart): a' = psi, u' = 0.  a' is shifted to max 0
// (the shift goes into the class offsets); M0' = sum exp(a'), Ma' = sum exp(a') a'.  All CTAs compute identical bits.
struct RcgsStep { double m0n, man, shift, beta; };
template <int MODE, int NT>
__device__ __forceinline__ RcgsStep rcgs_prep_step(const ViArrays &va, const RcgsGroup &g, const ViCtl *ctl, int K, double *s_an,
                                                   double *s_En, double *scratch) {
  double beta_eff = 0.0;
  if (MODE == 0) {
    const double beta = va.red[K + RS_NORM] / ctl->oldnorm;
    // the direction memory is empty before the first accepted step and after a restart
    beta_eff = (!ctl->didreset && beta > 0.0 && ctl->iter > 0) ? beta : 0.0;
  }
  const bool publish = blockIdx.x == 0;
  double mx = -INFINITY;
  for (int k = threadIdx.x; k < K; k += NT) {
    double an, un;
    if (MODE == 0) { un = fma(beta_eff, g.u[k], va.dg[k] - g.a[k]); an = g.a[k] + un; }
    else { un = 0.0; an = va.dg[k]; }
    if (publish) g.u_new[k] = un;
    s_an[k] = an;
    mx = fmax(mx, an);
  }
  mx = block_max<NT>(mx, scratch);
  double m0 = 0.0, ma = 0.0;
  for (int k = threadIdx.x; k < K; k += NT) {
    const double an = s_an[k] - mx, E = exp(an);
    s_an[k] = an; s_En[k] = E;
    if (publish) g.a_new[k] = an;
    m0 += E; ma = fma(E, an, ma);
  }
  m0 = block_sum<NT>(m0, scratch);
  ma = block_sum<NT>(ma, scratch);
  return RcgsStep{m0, ma, mx, beta_eff};
}

// ---- warp-level plumbing shared by the two sweeps ------------------------------------------------------------------
constexpr int RS_NT = 256;
constexpr int RS_STAGE = 224;      // hits a warp parks at a time
constexpr int RS_PF = 4;           // sweep A: slabs of 32 hits of the NEXT chunk kept in flight in registers
constexpr int RS_STAGE_B = 192;    // sweep B parks three doubles per hit (gamma, logl, exp(gamma - class max)): two CTAs per SM up to K = 2000
constexpr int RS_BATCH = 3;        // sweep B: slabs of the current chunk loaded together behind the prefetched one
constexpr int RS_PFB = 1;          // sweep B: one slab (four values per hit: deeper prefetch spills at two CTAs per SM, and one CTA
                                   // per SM with four slabs measured slower: 2.34 vs 1.93 ms on 2e7 x 2000)

// Sweep A: newnorm = sum_j (S2_j - S1_j^2), the direction centred per class by c_j = (l0 - b_j) + ebar so that every term
// vanishes at the optimum (no cancellation between large sums):
//   delta_e = (logl_e + psi_k - g_e) - c_j   on the hits,   ec_k = e_k - ebar on the non-hit groups (sum_k E_k ec_k = 0)
//   S1_j = sum_hits (q_e delta_e - exp(b_j) E_k ec_k)
//   S2_j = exp(b_j) M2c + sum_hits (q_e delta_e^2 - exp(b_j) E_k ec_k^2)
// (the body is shared with rcgs_fused_kernel; it ends with the CTA's partial norm stored)
__device__ __forceinline__ void
rcgs_sweep_a_body(const uint64_t *__restrict__ nz_ptr, const uint32_t *__restrict__ nz_grp, const double *__restrict__ nz_logl,
                  const double *sp_b, const double *sp_g, const ViArrays &va, const RcgsGroup &grp,
                  double *partials, int pstride, unsigned long long N, unsigned long long nnz, int K, double l0) {
  extern __shared__ __align__(16) unsigned char s_dyn_rs[];
  double *s_E = reinterpret_cast<double *>(s_dyn_rs);            // [K] exp(a_k)
  double *s_ec = s_E + K;                                         // [K]
  double *s_psi = s_ec + K;                                       // [K]
  double *s_x1 = s_psi + K;                                       // [warps][RS_STAGE]
  double *s_x2 = s_x1 + (RS_NT / 32) * RS_STAGE;                  // [warps][RS_STAGE]
  double *s_cls = s_x2 + (RS_NT / 32) * RS_STAGE;                 // [warps][2][32]: exp(b_j), centre c_j
  __shared__ double s_blk[32];
  for (int k = threadIdx.x; k < K; k += RS_NT) { s_E[k] = exp(grp.a[k]); s_ec[k] = grp.ec[k]; s_psi[k] = va.dg[k]; }
  const double ebar = grp.mom[RM_EBAR], m2c = grp.mom[RM_M2C];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double *x1 = s_x1 + warp * RS_STAGE, *x2 = s_x2 + warp * RS_STAGE;
  double *cls_w = s_cls + warp * 64, *cls_c = cls_w + 32;
  double nn = 0.0;
  const unsigned long long n_chunks = (N + 31) / 32;
  const unsigned long long warps_total = (unsigned long long)gridDim.x * (RS_NT / 32);
  const unsigned long long per = (n_chunks + warps_total - 1) / warps_total;
  const unsigned long long ch_begin = min(n_chunks, ((unsigned long long)blockIdx.x * (RS_NT / 32) + warp) * per);
  const unsigned long long ch_end = min(n_chunks, ch_begin + per);
  // A warp walks a contiguous run of chunks: while chunk i is being summed, the per-class values of chunk i + 1 and the
  // first RS_PF x 32 hits behind the current range are already in flight in registers (the sweep is bound by load latency).
  const unsigned long long last_hit = nnz ? nnz - 1 : 0;
  struct ClsA { unsigned long long a, b; double bj; bool have; };
  auto load_cls = [&](unsigned long long ch) {
    const unsigned long long j = ch * 32 + lane;
    ClsA c;
    c.have = j < N;
    c.a = nz_ptr[c.have ? j : N]; c.b = nz_ptr[c.have ? j + 1 : N];
    c.bj = c.have ? sp_b[j] : 0.0;
    return c;
  };
  uint32_t pk[RS_PF];
  double pl[RS_PF], pg[RS_PF];
  auto load_hits = [&](unsigned long long base) {
#pragma unroll
    for (int k = 0; k < RS_PF; ++k) {
      const unsigned long long e = min(base + (unsigned long long)(32 * k + lane), last_hit);
      pk[k] = nz_grp[e]; pl[k] = nz_logl[e]; pg[k] = sp_g[e];
    }
  };
  auto hit_terms = [&](uint32_t kg, double lg, double g, double &t1, double &t2) {
    const int k = (int)(kg & SP_GRP_MASK), c = (int)(kg >> 24);
    const double delta = (lg + s_psi[k] - g) - cls_c[c];
    const double q = exp_nonpos(fmin(g, 0.0));
    const double Ew = cls_w[c] * s_E[k], ec = s_ec[k];
    t1 = fma(q, delta, -Ew * ec);
    t2 = fma(q * delta, delta, -Ew * ec * ec);
  };
  if (ch_begin < ch_end) {
    ClsA nxt = load_cls(ch_begin);
    load_hits(__shfl_sync(0xffffffffu, nxt.a, 0));
    for (unsigned long long ch = ch_begin; ch < ch_end; ++ch) {
      const ClsA cur = nxt;
      const unsigned long long h0 = __shfl_sync(0xffffffffu, cur.a, 0), h1 = __shfl_sync(0xffffffffu, cur.b, 31);
      if (ch + 1 < ch_end) nxt = load_cls(ch + 1);
      const double wj = cur.have ? exp(cur.bj) : 0.0;
      __syncwarp();
      cls_w[lane] = wj;
      cls_c[lane] = (l0 - cur.bj) + ebar;
      __syncwarp();
      double s1 = 0.0, s2 = wj * m2c;
      if (h1 - h0 <= RS_STAGE) {
        const int n_here = (int)(h1 - h0);
#pragma unroll
        for (int k = 0; k < RS_PF; ++k) {
          const int x = 32 * k + lane;
          if (x < n_here) hit_terms(pk[k], pl[k], pg[k], x1[x], x2[x]);
        }
        for (int x = 32 * RS_PF + lane; x < n_here; x += 32) hit_terms(nz_grp[h0 + x], nz_logl[h0 + x], sp_g[h0 + x], x1[x], x2[x]);
        if (ch + 1 < ch_end) load_hits(h1);
        __syncwarp();
        for (int x = (int)(cur.a - h0), xe = (int)(cur.b - h0); x < xe; ++x) { s1 += x1[x]; s2 += x2[x]; }   // class-parallel, list order
      } else {
        for (unsigned long long q0 = h0; q0 < h1; q0 += RS_STAGE) {
          const int n_here = (int)min((unsigned long long)RS_STAGE, h1 - q0);
          __syncwarp();
          for (int x = lane; x < n_here; x += 32) hit_terms(nz_grp[q0 + x], nz_logl[q0 + x], sp_g[q0 + x], x1[x], x2[x]);
          __syncwarp();
          const unsigned long long lo = max(cur.a, q0), hi = min(cur.b, q0 + (unsigned long long)n_here);
          for (unsigned long long x = lo; x < hi; ++x) { s1 += x1[x - q0]; s2 += x2[x - q0]; }
        }
        if (ch + 1 < ch_end) load_hits(h1);
      }
      if (cur.have) nn += s2 - s1 * s1;
    }
  }
  nn = block_sum<RS_NT>(nn, s_blk);
  if (threadIdx.x == 0) partials[(unsigned long long)blockIdx.x * pstride + K + RS_NORM] = nn;
}
// the gradient norm from the CTAs' partial norms (one CTA, fixed order)
__device__ __forceinline__ void rcgs_sum_norms(const double *partials, int pstride, const ViArrays &va, int K) {
  __shared__ double s_blk[32];
  double acc = 0.0;
  for (int c = threadIdx.x; c < (int)gridDim.x; c += RS_NT) acc += __ldcg(partials + (size_t)c * pstride + K + RS_NORM);
  acc = block_sum<RS_NT>(acc, s_blk);
  if (threadIdx.x == 0) va.red[K + RS_NORM] = acc;
}
static __global__ void __launch_bounds__(RS_NT, 2)
rcgs_sweep_a_kernel(const uint64_t *__restrict__ nz_ptr, const uint32_t *__restrict__ nz_grp, const double *__restrict__ nz_logl,
                    const double *__restrict__ sp_b, const double *__restrict__ sp_g, ViArrays va, RcgsGroup grp, ViCtl *ctl,
                    double *partials, int pstride, unsigned long long N, unsigned long long nnz, int K, double l0) {
  if (ctl->done || ctl->stall) return;
  rcgs_sweep_a_body(nz_ptr, nz_grp, nz_logl, sp_b, sp_g, va, grp, partials, pstride, N, nnz, K, l0);
  if (cta_is_last(ctl)) {
    rcgs_sum_norms(partials, pstride, va, K);
    if (threadIdx.x == 0) ctl->ticket = 0;
  }
}
inline size_t rcgs_sweep_a_smem(int K) { return (size_t)K * 24 + (size_t)(RS_NT / 32) * RS_STAGE * 16 + (size_t)(RS_NT / 32) * 64 * 8; }

// Control step after sweep B (stage 0) / the restart sweep (stage 1); all NT threads of one CTA.
//   red[k] = sum over the hits of k of c_j (q_jk - exp(b_j) E'_k),  red[K + RS_MASS] = B = sum_j c_j exp(b_j),
//   red[K + RS_BOUND] = data term of the bound, red[K + RS_NORM] = the gradient norm of sweep A.
// Accepting commits the candidate group vectors and prepares the moments of the next sweep A.
template <int NT>
__device__ __forceinline__ void rcgs_ctl_b_step(const ViArrays &va, const RcgsGroup &g, ViCtl *ctl, int K, int stage,
                                                int stall_on_reject, double *scratch) {
  __shared__ int s_accept;
  const double mass = __ldcg(va.red + K + RS_MASS);
  double lg = 0.0;
  for (int k = threadIdx.x; k < K; k += NT) lg += lgamma(va.alpha0[k] + fma(exp(__ldcg(g.a_new + k)), mass, __ldcg(va.red + k)));
  lg = block_sum<NT>(lg, scratch);
  const double cand = __ldcg(va.red + K + RS_BOUND) + lg + ctl->bound_const;
  if (threadIdx.x == 0) {
    if (stage == 0) {
      const double nn = __ldcg(va.red + K + RS_NORM);
      ctl->beta = nn / ctl->oldnorm;
      ctl->newnorm = nn;
      ctl->oldnorm = nn;
      ctl->didreset = 0;
    }
    if (stage == 0 && cand < ctl->bound) {
      ctl->didreset = 1;
      ctl->resets += 1;
      if (stall_on_reject) ctl->stall = 1;
      s_accept = 0;
    } else {
      s_accept = 1;
      ctl->oldbound = ctl->bound;
      ctl->bound = cand;
      trace_push(va, ctl, ctl->newnorm, stage);
      ctl->iter += 1;
      ctl->stall = 0;
      if (stage == 0 && cand - ctl->oldbound < ctl->tol) ctl->converged = 1;
      if (ctl->converged || ctl->iter >= ctl->max_iters) ctl->done = 1;
    }
  }
  __syncthreads();
  if (!s_accept) return;
  for (int k = threadIdx.x; k < K; k += NT) {
    const double an = __ldcg(g.a_new + k);
    const double nk = va.alpha0[k] + fma(exp(an), mass, __ldcg(va.red + k));
    va.N_k[k] = nk;
    va.dg[k] = digamma_series(nk) - 1.0;
    g.a[k] = an;
    g.u[k] = __ldcg(g.u_new + k);
  }
  __syncthreads();
  rcgs_prep_a<NT>(va, g, K, scratch);
}

static __global__ void __launch_bounds__(256) rcgs_ctl_b_kernel(ViArrays va, RcgsGroup g, ViCtl *ctl, int K, int stage, int stall_on_reject) {
  if (ctl->done) return;
  if (stage == 0 ? ctl->stall != 0 : !ctl->didreset) return;
  __shared__ double scratch[32];
  rcgs_ctl_b_step<256>(va, g, ctl, K, stage, stall_on_reject, scratch);
}

// Sweep B.  MODE 0: the RCG step; MODE 1: the restart, gamma = normalise(logl + psi) (class part b = l0 + shift, hits
// logl_e + psi_k), no direction memory.  Per class j with the candidate group vectors (a', E' = exp(a'), M0', Ma', shift):
//   v_j' = (l0 - b_j) + beta v_j,  bb = b_j + v_j' + shift      (non-hit gamma before normalisation: a'_k + bb)
//   t_e' = (logl_e + psi_k - g_e) + beta t_e,  gg_e = g_e + t_e'
//   m_j  = log( exp(bb) (M0' - sum_H E'_k) + sum_hits exp(gg_e) ),   b_j' = bb - m_j,  g_e' = gg_e - m_j
// Three rounds over the chunk's hits: (0) new step out, gamma parked, class maximum and hit-group moments; (1) the
// normaliser; (2) normalised gamma out, N_k scatter, bound.  A chunk that fits one stage is parked once.
// tail: 0 none (rcgs_finalize_kernel follows), 1 the last CTA reduces the partial vectors, 2 ... and takes the control step.
// (the body is shared with rcgs_fused_kernel; it ends with the CTA's partial vector stored)
template <int MODE>
__device__ __forceinline__ void
rcgs_sweep_b_body(const uint64_t *__restrict__ nz_ptr, const uint32_t *__restrict__ nz_grp, const double *__restrict__ nz_logl,
                  const double *__restrict__ counts, double *sp_b, double *sp_v, double *sp_g, double *sp_t,
                  const ViArrays &va, const RcgsGroup &grp, const ViCtl *ctl,
                  double *partials, int pstride, unsigned long long N, unsigned long long nnz, int K, double l0, double fx_scale) {
  extern __shared__ __align__(16) unsigned char s_dyn_rs[];
  double *s_an = reinterpret_cast<double *>(s_dyn_rs);            // [K] a'_k
  double *s_En = s_an + K;                                         // [K] exp(a'_k)
  double *s_psi = s_En + K;                                        // [K]
  unsigned *s_acc = reinterpret_cast<unsigned *>(s_psi + K);       // [K][2] fixed-point accumulators
  double *s_gg = reinterpret_cast<double *>(s_acc + 2 * (size_t)K);   // [warps][RS_STAGE_B] gamma of the parked hits (before normalisation)
  double *s_ll = s_gg + (RS_NT / 32) * RS_STAGE_B;                 // [warps][RS_STAGE_B] logl of the parked hits
  double *s_ee = s_ll + (RS_NT / 32) * RS_STAGE_B;                 // [warps][RS_STAGE_B] exp(gamma - class maximum)
  double *s_cls = s_ee + (RS_NT / 32) * RS_STAGE_B;                // [warps][3][32]: per class of the chunk
  uint32_t *s_key = reinterpret_cast<uint32_t *>(s_cls + (RS_NT / 32) * 96);   // [warps][RS_STAGE_B]
  __shared__ double s_blk[32];
  for (int k = threadIdx.x; k < K; k += RS_NT) { s_psi[k] = va.dg[k]; s_acc[2 * k] = 0u; s_acc[2 * k + 1] = 0u; }
  const RcgsStep stp = rcgs_prep_step<MODE, RS_NT>(va, grp, ctl, K, s_an, s_En, s_blk);
  const double m0n = stp.m0n, man = stp.man, shift = stp.shift;
  const double beta = MODE == 0 ? stp.beta : 0.0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double *gg = s_gg + warp * RS_STAGE_B, *ll = s_ll + warp * RS_STAGE_B, *ee = s_ee + warp * RS_STAGE_B;
  uint32_t *key = s_key + warp * RS_STAGE_B;
  // m_j (the class maximum while the exponentials are taken), c_j exp(b_j'), c_j (one-piece chunks: c_j / normaliser)
  double *cls_m = s_cls + warp * 96, *cls_cw = cls_m + 32, *cls_c = cls_m + 64;
  double bound = 0.0, mass = 0.0;
  const unsigned long long n_chunks = (N + 31) / 32;
  const unsigned long long warps_total = (unsigned long long)gridDim.x * (RS_NT / 32);
  const unsigned long long per = (n_chunks + warps_total - 1) / warps_total;
  const unsigned long long ch_begin = min(n_chunks, ((unsigned long long)blockIdx.x * (RS_NT / 32) + warp) * per);
  const unsigned long long ch_end = min(n_chunks, ch_begin + per);
  // As in sweep A: the per-class values of chunk i + 1 and the first RS_PFB x 32 hits behind the current range are in
  // flight in registers while chunk i is processed (a hit of the next chunk is touched by no other warp meanwhile).
  const unsigned long long last_hit = nnz ? nnz - 1 : 0;
  struct ClsB { unsigned long long a, b; double c, bj, vj; bool have; };
  auto load_cls = [&](unsigned long long ch) {
    const unsigned long long j = ch * 32 + lane;
    ClsB d;
    d.have = j < N;
    d.a = nz_ptr[d.have ? j : N]; d.b = nz_ptr[d.have ? j + 1 : N];
    d.c = d.have ? counts[j] : 0.0;
    d.bj = (MODE == 0 && d.have) ? sp_b[j] : 0.0;
    d.vj = (MODE == 0 && d.have) ? sp_v[j] : 0.0;
    return d;
  };
  uint32_t pk[RS_PFB];
  double pl[RS_PFB], pg[RS_PFB], pt[RS_PFB];
  auto load_hits = [&](unsigned long long base) {
#pragma unroll
    for (int k = 0; k < RS_PFB; ++k) {
      const unsigned long long e = min(base + (unsigned long long)(32 * k + lane), last_hit);
      pk[k] = nz_grp[e]; pl[k] = nz_logl[e];
      if (MODE == 0) { pg[k] = sp_g[e]; pt[k] = sp_t[e]; }
    }
  };
  // round 0 of one hit: the new direction leaves at once; returns gamma before normalisation
  auto step_hit = [&](unsigned long long e, uint32_t kg, double lg, double g, double t) {
    const int k = (int)(kg & SP_GRP_MASK);
    if (MODE == 0) {
      const double tn = fma(beta, t, lg + s_psi[k] - g);
      sp_t[e] = tn;
      return g + tn;
    }
    return lg + s_psi[k];
  };
  if (ch_begin < ch_end) {
    ClsB nxt = load_cls(ch_begin);
    load_hits(__shfl_sync(0xffffffffu, nxt.a, 0));
    for (unsigned long long ch = ch_begin; ch < ch_end; ++ch) {
      const ClsB cur = nxt;
      const unsigned long long j = ch * 32 + lane;
      const unsigned long long h0 = __shfl_sync(0xffffffffu, cur.a, 0), h1 = __shfl_sync(0xffffffffu, cur.b, 31);
      if (ch + 1 < ch_end) nxt = load_cls(ch + 1);
      const double vn = MODE == 0 ? fma(beta, cur.vj, l0 - cur.bj) : 0.0;     // class part of the step
      const double bb = MODE == 0 ? cur.bj + vn + shift : l0 + shift;          // class part of gamma before normalisation
      const bool one_piece = h1 - h0 <= RS_STAGE_B;
      double mx = bb, sumE = 0.0, sumEa = 0.0, ssum = 0.0;
      // normaliser of the class: exp(bb - mx) (M0' - sum_H E') + sum_hits exp(gg - mx); leaves m_j, c_j exp(b_j') and c_j
      // (scaled: c_j / normaliser, so that c_j q = c_j exp(gg - mx) / normaliser needs no second exponential per hit)
      auto class_normaliser = [&](bool scaled) {
        const double rest = fmax(m0n - sumE, 0.0);
        const double tot = fma(exp_nonpos(bb - mx), rest, ssum);
        const double m = mx + log(tot);
        const double bnew = bb - m;
        const double wnew = exp(bnew);
        if (cur.have) {
          sp_b[j] = bnew;
          if (MODE == 0) sp_v[j] = vn;
          if (cur.c > 0.0) {
            mass = fma(cur.c, wnew, mass);
            // non-hit part of the bound: c_j exp(b_j') ((l0 - b_j') (M0' - sum_H E') - (Ma' - sum_H E' a'))
            bound = fma(cur.c * wnew, (l0 - bnew) * rest - (man - sumEa), bound);
          }
        }
        __syncwarp();
        cls_m[lane] = m;
        cls_cw[lane] = cur.c * wnew;
        cls_c[lane] = scaled ? cur.c / tot : cur.c;
        __syncwarp();
      };
      if (one_piece) {
        // The chunk's hits are parked once.  ONE exponential per hit, taken hit-parallel (every lane busy) against the class
        // maximum; the class sums it, and c_j q(j,k) = exp(gg - mx) c_j / normaliser.
        const int n_here = (int)(h1 - h0);
        __syncwarp();
#pragma unroll
        for (int k = 0; k < RS_PFB; ++k) {                        // the hits that arrived in registers
          const int x = 32 * k + lane;
          if (x < n_here) { key[x] = pk[k]; ll[x] = pl[k]; gg[x] = step_hit(h0 + x, pk[k], pl[k], pg[k], pt[k]); }
        }
        // the rest straight from memory, three slabs at a time with all their loads issued before the first use (one exposed
        // round trip for a typical chunk of four to five slabs instead of one per slab: ncu had these loads as the top stall)
        for (int x0 = 32 * RS_PFB + lane; x0 < n_here; x0 += 32 * RS_BATCH) {
          uint32_t kg[RS_BATCH];
          double lg[RS_BATCH], g0[RS_BATCH], t0[RS_BATCH];
#pragma unroll
          for (int u = 0; u < RS_BATCH; ++u) {
            const unsigned long long e = h0 + (unsigned long long)min(x0 + 32 * u, n_here - 1);   // (clamped: loads carry no predicate)
            kg[u] = nz_grp[e]; lg[u] = nz_logl[e];
            g0[u] = MODE == 0 ? sp_g[e] : 0.0; t0[u] = MODE == 0 ? sp_t[e] : 0.0;
          }
#pragma unroll
          for (int u = 0; u < RS_BATCH; ++u) {
            const int x = x0 + 32 * u;
            if (x < n_here) { key[x] = kg[u]; ll[x] = lg[u]; gg[x] = step_hit(h0 + x, kg[u], lg[u], g0[u], t0[u]); }
          }
        }
        if (ch + 1 < ch_end) load_hits(h1);
        __syncwarp();
        const int xa = (int)(cur.a - h0), xb = (int)(cur.b - h0);
        for (int x = xa; x < xb; ++x) {                           // class-parallel, list order
          mx = fmax(mx, gg[x]);
          const int k = (int)(key[x] & SP_GRP_MASK);
          sumE += s_En[k];
          sumEa = fma(s_En[k], s_an[k], sumEa);
        }
        __syncwarp();
        cls_m[lane] = mx;
        __syncwarp();
        for (int x = lane; x < n_here; x += 32) ee[x] = exp_nonpos(gg[x] - cls_m[key[x] >> 24]);   // hit-parallel
        __syncwarp();
        for (int x = xa; x < xb; ++x) ssum += ee[x];              // class-parallel, list order
        class_normaliser(true);
        for (int x = lane; x < n_here; x += 32) {                 // hit-parallel: normalised gamma out, N_k scatter, bound
          const uint32_t kg = key[x];
          const int k = (int)(kg & SP_GRP_MASK), cl = (int)(kg >> 24);
          const double gnew = gg[x] - cls_m[cl];
          sp_g[h0 + x] = gnew;
          const double ct = cls_c[cl];
          if (ct > 0.0) {
            const double cq = ct * ee[x];                                      // c_j q(j,k)
            bound = fma(cq, ll[x] - gnew, bound);
            const double val = cq - cls_cw[cl] * s_En[k];                      // the group's closed-form share already counts exp(b_j') E'_k
            if (val != 0.0) fx_atomic_add(&s_acc[2 * k], __double2ll_rn(val * fx_scale));
          }
        }
      } else {
        // a chunk in several pieces: parked again in every round (0: new step out, class maximum and hit-group moments;
        // 1: the normaliser; 2: normalised gamma out, N_k scatter, bound)
        for (int round = 0; round < 3; ++round) {
          for (unsigned long long q0 = h0; q0 < h1; q0 += RS_STAGE_B) {
            const int n_here = (int)min((unsigned long long)RS_STAGE_B, h1 - q0);
            __syncwarp();
            for (int x = lane; x < n_here; x += 32) {
              const uint32_t kg = nz_grp[q0 + x];
              const double lg = nz_logl[q0 + x];
              double g2;
              if (MODE == 0) {
                const double g = sp_g[q0 + x];
                g2 = round == 0 ? step_hit(q0 + x, kg, lg, g, sp_t[q0 + x]) : g + sp_t[q0 + x];   // (later rounds: the step is already the new one)
              } else {
                g2 = lg + s_psi[kg & SP_GRP_MASK];
              }
              key[x] = kg; gg[x] = g2; ll[x] = lg;
            }
            __syncwarp();
            const unsigned long long lo = max(cur.a, q0), hi = min(cur.b, q0 + (unsigned long long)n_here);
            if (round == 0) {                                         // class-parallel, list order
              for (unsigned long long x = lo; x < hi; ++x) {
                mx = fmax(mx, gg[x - q0]);
                const int k = (int)(key[x - q0] & SP_GRP_MASK);
                sumE += s_En[k];
                sumEa = fma(s_En[k], s_an[k], sumEa);
              }
            } else if (round == 1) {
              for (unsigned long long x = lo; x < hi; ++x) ssum += exp_nonpos(gg[x - q0] - mx);
            } else {
              for (int x = lane; x < n_here; x += 32) {               // hit-parallel
                const uint32_t kg = key[x];
                const int k = (int)(kg & SP_GRP_MASK), cl = (int)(kg >> 24);
                const double gnew = gg[x] - cls_m[cl];
                sp_g[q0 + x] = gnew;
                const double cj = cls_c[cl];
                if (cj > 0.0) {
                  const double cq = cj * exp_nonpos(fmin(gnew, 0.0));           // c_j q(j,k)
                  bound = fma(cq, ll[x] - gnew, bound);
                  const double val = cq - cls_cw[cl] * s_En[k];
                  if (val != 0.0) fx_atomic_add(&s_acc[2 * k], __double2ll_rn(val * fx_scale));
                }
              }
            }
          }
          if (round == 0 && ch + 1 < ch_end) load_hits(h1);
          if (round == 1) class_normaliser(false);
        }
      }
    }
  }
  __syncthreads();
  double *out = partials + (unsigned long long)blockIdx.x * pstride;
  const double fx_inv = 1.0 / fx_scale;
  for (int k = threadIdx.x; k < K; k += RS_NT) {
    const long long v = (long long)(((unsigned long long)s_acc[2 * k + 1] << 32) | (unsigned long long)s_acc[2 * k]);
    out[k] = (double)v * fx_inv;
  }
  bound = block_sum<RS_NT>(bound, s_blk);
  mass = block_sum<RS_NT>(mass, s_blk);
  if (threadIdx.x == 0) { out[K + RS_BOUND] = bound; out[K + RS_MASS] = mass; }
}
template <int MODE, bool TAIL>
__global__ void __launch_bounds__(RS_NT, 2)
rcgs_sweep_b_kernel(const uint64_t *__restrict__ nz_ptr, const uint32_t *__restrict__ nz_grp, const double *__restrict__ nz_logl,
                    const double *__restrict__ counts, double *__restrict__ sp_b, double *__restrict__ sp_v,
                    double *__restrict__ sp_g, double *__restrict__ sp_t, ViArrays va, RcgsGroup grp, ViCtl *ctl,
                    double *partials, int pstride, unsigned long long N, unsigned long long nnz, int K, double l0, double fx_scale,
                    int tail) {
  if (ctl->done) return;
  if (MODE == 1 ? !ctl->didreset : ctl->stall != 0) return;
  rcgs_sweep_b_body<MODE>(nz_ptr, nz_grp, nz_logl, counts, sp_b, sp_v, sp_g, sp_t, va, grp, ctl, partials, pstride, N, nnz, K, l0, fx_scale);
  if constexpr (TAIL) {
    if (tail != 0 && cta_is_last(ctl)) {
      __shared__ double s_tail[32];
      reduce_partials_cta<RS_NT>(partials, pstride, (int)gridDim.x, K + 2, va.red, va.seg);   // (slot RS_NORM stays: sweep A's norm)
      if (threadIdx.x == 0) ctl->ticket = 0;
      if (tail == 2) {
        __syncthreads();
        rcgs_ctl_b_step<RS_NT>(va, grp, ctl, K, MODE, 0, s_tail);
      }
    }
  }
}

// ---- small problems: whole iterations in ONE launch ---------------------------------------------------------------
// When an iteration takes tens of microseconds, three launches and their drains are most of it.  rcgs_fused_kernel is a
// cooperative launch (every CTA resident) that runs up to `n_steps` iterations: sweep A, a grid rendezvous whose last
// arrival sums the norms, sweep B, a rendezvous whose last arrival reduces the partial vectors and takes the control step,
// and the restart sweep when that step was rejected.  One GPU only.  Used up to ~1 ms of sweeps per iteration (config 2 / 5:
// 0.17 ms of sweeps, 0.25 ms per iteration as four launches).
// (grid_rendezvous: vi_kernels.cuh.)
// coop_reduce: the partial vectors are too many for one CTA (grid x columns > tail_max): after a rendezvous EVERY CTA sums
// its tiles of columns (reduce_partials_tiled), and the last arrival of a second rendezvous takes the control step.
static __global__ void __launch_bounds__(RS_NT, 2)
rcgs_fused_kernel(const uint64_t *__restrict__ nz_ptr, const uint32_t *__restrict__ nz_grp, const double *__restrict__ nz_logl,
                  const double *__restrict__ counts, double *sp_b, double *sp_v, double *sp_g, double *sp_t, ViArrays va, RcgsGroup grp,
                  ViCtl *ctl, double *partials, int pstride, unsigned long long N, unsigned long long nnz, int K, double l0,
                  double fx_scale, unsigned long long n_steps, int coop_reduce) {
  __shared__ double s_tail[32];
  __shared__ double s_tile[RS_NT];
  unsigned epoch = *reinterpret_cast<volatile unsigned *>(&ctl->epoch);
  auto reduce_and_control = [&](int stage) {
    if (coop_reduce) {
      grid_rendezvous(ctl, epoch, [] {});
      reduce_partials_tiled<RS_NT>(partials, pstride, (int)gridDim.x, K + 2, va.red, s_tile);
      grid_rendezvous(ctl, epoch, [&] { rcgs_ctl_b_step<RS_NT>(va, grp, ctl, K, stage, 0, s_tail); });
    } else {
      grid_rendezvous(ctl, epoch, [&] {
        reduce_partials_cta<RS_NT>(partials, pstride, (int)gridDim.x, K + 2, va.red, va.seg);
        __syncthreads();
        rcgs_ctl_b_step<RS_NT>(va, grp, ctl, K, stage, 0, s_tail);
      });
    }
  };
  for (unsigned long long it = 0; it < n_steps; ++it) {
    if (*reinterpret_cast<volatile int *>(&ctl->done)) break;          // (the same value in every CTA: read between rendezvous)
    rcgs_sweep_a_body(nz_ptr, nz_grp, nz_logl, sp_b, sp_g, va, grp, partials, pstride, N, nnz, K, l0);
    grid_rendezvous(ctl, epoch, [&] { rcgs_sum_norms(partials, pstride, va, K); });
    rcgs_sweep_b_body<0>(nz_ptr, nz_grp, nz_logl, counts, sp_b, sp_v, sp_g, sp_t, va, grp, ctl, partials, pstride, N, nnz, K, l0, fx_scale);
    reduce_and_control(0);
    if (*reinterpret_cast<volatile int *>(&ctl->didreset)) {
      rcgs_sweep_b_body<1>(nz_ptr, nz_grp, nz_logl, counts, sp_b, sp_v, sp_g, sp_t, va, grp, ctl, partials, pstride, N, nnz, K, l0, fx_scale);
      reduce_and_control(1);
    }
  }
}
inline size_t rcgs_sweep_b_smem(int K) {
  return (size_t)K * 32 + (size_t)(RS_NT / 32) * RS_STAGE_B * (24 + 4) + (size_t)(RS_NT / 32) * 96 * 8;
}

// Reduction of the partial vectors of sweep B over several CTAs (large grids x many groups); ctl_stage >= 0: the last
// CTA takes the control step (one GPU; several GPUs with peer memory: after exchanging the vector), -1: the NCCL all-reduce
// and rcgs_ctl_b_kernel follow.
static __global__ void __launch_bounds__(FIN_NT)
rcgs_finalize_kernel(const double *partials, int pstride, int n_ctas, ViArrays va, RcgsGroup grp, ViCtl *ctl, int K, int ctl_stage,
                     int restart, int peer, PeerView pv) {
  if (ctl->done) return;
  if (restart ? !ctl->didreset : ctl->stall != 0) return;
  __shared__ double s_blk[32];
  __shared__ double s_tile[FIN_NT];
  reduce_partials_tiled<FIN_NT>(partials, pstride, n_ctas, K + 2, va.red, s_tile);
  if (ctl_stage < 0) return;
  if (!cta_is_last(ctl)) return;
  if (threadIdx.x == 0) ctl->ticket = 0;
  __syncthreads();
  if (peer) {                                   // several GPUs, peer memory: exchange, then the control step (peer.cuh)
    if (!peer_allreduce_cta<FIN_NT>(va.red, K + 2, pv)) {
      if (threadIdx.x == 0) { ctl->fault = 2; ctl->done = 1; }
      return;
    }
  }
  rcgs_ctl_b_step<FIN_NT>(va, grp, ctl, K, ctl_stage, peer ? 1 : 0, s_blk);
}

// Start: gamma = log(1/K) everywhere (a = 0, b = log(1/K), hits log(1/K)), no direction; moments for the first sweep A.
static __global__ void __launch_bounds__(256) rcgs_init_groups_kernel(ViArrays va, RcgsGroup g, int K) {
  __shared__ double scratch[32];
  for (int k = threadIdx.x; k < K; k += 256) { g.a[k] = 0.0; g.u[k] = 0.0; g.a_new[k] = 0.0; g.u_new[k] = 0.0; }
  __syncthreads();
  rcgs_prep_a<256>(va, g, K, scratch);
}

} // namespace mswb
