// vi_kernels.cuh — the fused VI sweeps (K6 / K7 of SURVEY.md §2.4), EC-major, HBM-bound.
//
// Every sweep has the same shape.  A CTA of NT threads is split into G = NT/TPR row groups; a group of
// TPR threads owns R consecutive classes (rows) per batch and strides the K columns of a row in
// 128-bit pieces: thread t of the group holds columns VEC*(t + TPR*i) .. +VEC-1, i < KITER.  Per-class
// quantities (the logsumexp / the normaliser S_j / the mean step) are reductions ALONG a row:
// warp shuffles, plus one shared-memory hop when a row spans several warps.  Per-group quantities
// (the expected counts N_k) accumulate DOWN the rows in registers and leave the CTA once, as a
// per-CTA partial vector; the partial vectors are summed in CTA order (no floating-point atomics: results are
// bit-reproducible and identical however the CTAs were scheduled).  The grid is persistent (one CTA per SM, or
// a small multiple) and walks the row batches with a grid stride.
//
// Short rows (K <= 64) use TPR = 4 / 8 / 16: a warp then works on 32/TPR row groups at once, the reductions stay
// inside TPR-lane segments, and every lane owns one row's division and logarithm.
//
// Who sums the partial vectors (argument `tail` of every sweep):
//   0  a separate kernel (finalize_ctl_kernel) — large grids x many groups, where one CTA would take too long;
//   1  the last CTA of the sweep to finish (a ticket in the control block), nothing else — several GPUs: the
//      all-reduce and the control kernel follow;
//   2  the last CTA, which then also takes the control step (N_k, bound, convergence, next digamma vector) —
//      one GPU, small problems: an EM iteration is ONE launch, an RCG iteration two.
//
// How the rows reach the SM (template parameter PIPE; which one a sweep uses by default was measured, vi.cu):
//   PIPE = true   whole row batches are pulled into a ring of shared-memory stages by the TMA unit
//                 (cp.async.bulk global -> shared, completion on an mbarrier with expect_tx); one
//                 elected thread keeps STAGES-1 batches in flight while all warps consume the oldest.
//                 Bytes in flight per SM = (STAGES-1) x stage size, independent of register pressure.
//                 Sized for two resident CTAs it is the default of the RCG sweeps with 64-256 threads per row.
//   PIPE = false  direct 128-bit streaming loads into registers; the EM pass software-pipelines them one batch
//                 ahead (two register buffers).  Default of the EM pass and of the remaining RCG shapes.
#pragma once
#include "common.cuh"
#include "mathfn.cuh"
#include <type_traits>

namespace mswb {

// Device-resident control block of one optimisation: the host never has to be in the loop.
struct ViCtl {
  double bound, oldbound;
  double oldnorm, newnorm, beta;
  double bound_const, tol, sum_counts, dg_max;
  unsigned long long iter, max_iters, resets;
  int use_old, didreset, converged, done, fault;
  int stall;             // several GPUs: an RCG step lost ground; iterations pause until the host has enqueued the restart
  unsigned int ticket;   // CTAs of the running sweep that have delivered their partial vector (the last one reduces them)
  unsigned int epoch;    // grid rendezvous of the fused small-problem kernels: bumped by the last arrival
};

// Per-group vectors of one optimisation (all device pointers).
struct ViArrays {
  double *alpha0, *N_k, *dg, *w;   // [K]   dg = digamma(N_k) (EM) or digamma(N_k) - 1 (RCG)
  double *dg_prev;                  // [K]   EM: the digamma vector the LAST pass used (posteriors on demand)
  double *red;                      // [K + 3] reduced sums of the last sweep: per group, then the three slots below
  double *seg;                      // [RED_SEGS x pstride] scratch of the last CTA's segmented reduction
  double *trace_bound, *trace_gnorm;
  unsigned char *trace_reset;
  unsigned long long trace_cap;
};
// Slots of a partial / reduced vector behind the K per-group sums.
//   RED_BOUND  data term of the bound;  RED_AUX  EM sparse: the share every group receives, RCG: the gradient norm;
//   RED_FAULT  > 0 when a class normaliser left the representable range on ANY rank (summed by the all-reduce, so that
//              every rank takes the same decision).
constexpr int RED_BOUND = 0, RED_AUX = 1, RED_FAULT = 2, RED_EXTRA = 3;
constexpr int RED_SEGS = 8;

// TPR threads per row.  TPR >= 32: whole warps per row.  TPR = 4 / 8 / 16: several rows per warp (short rows).
template <int TPR_, int KITER_, int R_> struct Tile {
  static constexpr int TPR = TPR_, KITER = KITER_, R = R_;
  static constexpr int NT = TPR < 256 ? (256 / TPR) * TPR : TPR;   // as many whole row groups as fit in 256 threads
  static constexpr int G = NT / TPR;
  static constexpr int NW = NT / 32;
  static constexpr int WPG = TPR / 32;            // warps per row group (0: the row group is a segment of a warp)
  static constexpr int W = TPR < 32 ? TPR : 32;   // lanes of one warp that work on the same rows
  static_assert(TPR >= 32 || (32 % TPR == 0 && R <= TPR), "sub-warp row groups: TPR divides 32 and R <= TPR");
};

// Barrier over the TPR threads of one row group: a named barrier (bar.sync id, count) when the CTA holds several groups,
// so that groups drift apart instead of marching in CTA-wide lockstep; the CTA barrier when the row spans the CTA.
template <class TL> __device__ __forceinline__ void group_sync(int g) {
  if constexpr (TL::G == 1) __syncthreads();
  else asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(TL::TPR) : "memory");   // ids 1..G (G <= 8), 0 is __syncthreads
}

// ---- TMA bulk copy + mbarrier (inline PTX; SASS: UBLKCP / SYNCS) -------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// A ring of shared-memory stages; each stage holds `stage_rows` consecutive rows of up to NSRC arrays
// that share one row pitch.  Pipe batch p covers rows [p*stage_rows, (p+1)*stage_rows); CTA b takes
// p = b, b + gridDim, ...  All threads call run(); thread 0 is the producer.
template <int NSRC> struct RowPipe {
  unsigned char *buf;        // stages x NSRC x stage_pitch bytes
  uint64_t *full;            // one mbarrier per stage
  int stages, stage_rows, nsrc;
  uint32_t row_bytes, stage_pitch;
  const unsigned char *src[NSRC];
  unsigned long long N;

  __device__ __forceinline__ void issue(int stage, unsigned long long p) const {
    const unsigned long long row0 = p * (unsigned long long)stage_rows;
    const unsigned long long rows = min((unsigned long long)stage_rows, N - row0);
    const uint32_t bytes = (uint32_t)rows * row_bytes;
    mbar_expect_tx(&full[stage], bytes * (uint32_t)nsrc);
    constexpr uint32_t CHUNK = 32768;     // several bulk copies per stage keep more of the TMA unit busy
#pragma unroll
    for (int s = 0; s < NSRC; ++s) {
      if (s >= nsrc) break;
      unsigned char *dst = buf + ((size_t)stage * NSRC + s) * stage_pitch;
      const unsigned char *from = src[s] + row0 * (unsigned long long)row_bytes;
      for (uint32_t o = 0; o < bytes; o += CHUNK) bulk_g2s(dst + o, from + o, min(CHUNK, bytes - o), &full[stage]);
    }
  }
  __device__ __forceinline__ const unsigned char *stage_ptr(int stage, int s) const {
    return buf + ((size_t)stage * NSRC + s) * stage_pitch;
  }

  // consume(stage, first_row_of_stage) is called by all threads once per pipe batch, in order.
  template <class F> __device__ __forceinline__ void run(F &&consume) {
    if (threadIdx.x == 0) {
      for (int s = 0; s < stages; ++s) mbar_init(&full[s], 1);
      mbar_fence_init();
    }
    __syncthreads();
    const unsigned long long n_pb = (N + stage_rows - 1) / stage_rows;
    const unsigned long long mine = n_pb > blockIdx.x ? (n_pb - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    if (threadIdx.x == 0) {
      const unsigned long long pre = min((unsigned long long)(stages - 1), mine);
      for (unsigned long long i = 0; i < pre; ++i) issue((int)i, blockIdx.x + i * gridDim.x);
    }
    int stage = 0;
    uint32_t parity = 0;
    for (unsigned long long i = 0; i < mine; ++i) {
      __syncthreads();                         // every warp is done with the stage consumed last time round
      if (threadIdx.x == 0) {
        const unsigned long long nxt = i + (unsigned long long)(stages - 1);
        if (nxt < mine) issue((int)(nxt % (unsigned long long)stages), blockIdx.x + nxt * gridDim.x);
      }
      mbar_wait(&full[stage], parity);
      consume(stage, (blockIdx.x + i * gridDim.x) * (unsigned long long)stage_rows);
      if (++stage == stages) { stage = 0; parity ^= 1; }
    }
  }
};

// ---- per-CTA partial vector: fold the row groups in a fixed order, then one coalesced store ----------
// acc[i][v] belongs to column VEC*(t + TPR*i) + v.  comb: TPR*KITER*VEC doubles when G > 1.
template <class TL, int VEC>
__device__ __forceinline__ void store_partials(const double (&acc_in)[TL::KITER][VEC], double *comb, double *out, int K) {
  const int t = threadIdx.x % TL::TPR, g = threadIdx.x / TL::TPR;
  if constexpr (TL::G == 1) {
#pragma unroll
    for (int i = 0; i < TL::KITER; ++i)
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const int k = VEC * (t + TL::TPR * i) + v;
        if (k < K) out[k] = acc_in[i][v];
      }
  } else if constexpr (TL::TPR < 32) {
    // the row groups of a warp hold the same columns TPR lanes apart: butterfly over them first (fixed order), then
    // the warps in warp order through shared memory
    double acc[TL::KITER][VEC];
#pragma unroll
    for (int i = 0; i < TL::KITER; ++i)
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        double a = acc_in[i][v];
#pragma unroll
        for (int off = 16; off >= TL::TPR; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
        acc[i][v] = a;
      }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int ww = 0; ww < TL::NW; ++ww) {
      __syncthreads();
      if (warp == ww && lane < TL::TPR) {
#pragma unroll
        for (int i = 0; i < TL::KITER; ++i)
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            const int c = VEC * (lane + TL::TPR * i) + v;
            comb[c] = ww == 0 ? acc[i][v] : comb[c] + acc[i][v];
          }
      }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += TL::NT) out[k] = comb[k];
  } else {
    for (int gg = 0; gg < TL::G; ++gg) {      // fixed order: bit-reproducible
      __syncthreads();
      if (g == gg) {
#pragma unroll
        for (int i = 0; i < TL::KITER; ++i)
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            const int c = VEC * (t + TL::TPR * i) + v;
            comb[c] = gg == 0 ? acc_in[i][v] : comb[c] + acc_in[i][v];
          }
      }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += TL::NT) out[k] = comb[k];
  }
}

template <typename ST> struct VecOf;
template <> struct VecOf<double> { using type = double2; static constexpr int VEC = 2; };
template <> struct VecOf<float> { using type = float4; static constexpr int VEC = 4; };

__device__ __forceinline__ void unpack(const double2 &x, double (&o)[2]) { o[0] = x.x; o[1] = x.y; }
__device__ __forceinline__ void unpack(const float4 &x, float (&o)[4]) { o[0] = x.x; o[1] = x.y; o[2] = x.z; o[3] = x.w; }

// Pipeline geometry chosen on the host (vi.cu) and passed to every PIPE kernel.
struct PipeGeom { int stages, stage_rows; unsigned stage_pitch; };

extern __shared__ __align__(128) unsigned char g_dyn_smem[];

// =====================================================================================================
// Reductions of R per-row values over the W lanes that share the rows (W = 32, or the TPR lanes of a sub-warp group).
// A transposing butterfly: while more than one row is left a lane keeps half of the rows and hands the other half
// to its partner (n/2 shuffles per step instead of n), then plain xor steps.  R - 1 + log2(W / R) 64-bit shuffles
// instead of R log2(W).  On return every lane holds the total of row row_of_lane(lane), W / R lanes per row.
// =====================================================================================================
template <bool IS_MAX> __device__ __forceinline__ double red_op(double a, double b) { return IS_MAX ? fmax(a, b) : a + b; }

template <int N, int OFF, bool IS_MAX> struct RowsRed {
  static __device__ __forceinline__ double run(const double (&v)[N], int lane) {
    if constexpr (OFF == 0) {
      static_assert(N == 1, "rows per batch must not exceed the lanes that share them");
      return v[0];
    } else if constexpr (N > 1) {
      const bool hi = (lane & OFF) != 0;
      double nxt[N / 2];
#pragma unroll
      for (int i = 0; i < N / 2; ++i) {
        const double keep = hi ? v[i + N / 2] : v[i], send = hi ? v[i] : v[i + N / 2];
        nxt[i] = red_op<IS_MAX>(keep, __shfl_xor_sync(0xffffffffu, send, OFF));
      }
      return RowsRed<N / 2, OFF / 2, IS_MAX>::run(nxt, lane);
    } else {
      const double nxt[1] = {red_op<IS_MAX>(v[0], __shfl_xor_sync(0xffffffffu, v[0], OFF))};
      return RowsRed<1, OFF / 2, IS_MAX>::run(nxt, lane);
    }
  }
};
// the row whose total a lane holds after RowsRed<R, W/2>: the lane's transposing bits (W/2, W/4, ...), MSB first
template <int W, int R> __device__ __forceinline__ int row_of_lane(int lane) {
  int rid = 0;
#pragma unroll
  for (int n = R, off = W / 2; n > 1; n >>= 1, off >>= 1) rid = (rid << 1) | ((lane & off) ? 1 : 0);
  return rid;
}
// lane (inside its W-lane segment) that canonically holds row r: its transposing bits spell r, the others are 0
template <int W, int R> __device__ __forceinline__ int holder_lane(int r) {
  int l = 0;
#pragma unroll
  for (int n = R, off = W / 2; n > 1; n >>= 1, off >>= 1) l |= (r & (n >> 1)) ? off : 0;
  return l;
}

// Reduce R per-row values over the TPR threads of a row group (sum or max).  On return the lanes with
// (lane % W) < R hold the group-wide result of row (lane % W); other lanes hold unspecified values.
// scratch: 2 * NW * R doubles (used only when a row spans several warps).
template <class TL, bool IS_MAX>
__device__ __forceinline__ double rows_reduce(const double (&v)[TL::R], double *scratch, int &phase, int lane, int warp) {
  constexpr int R = TL::R, W = TL::W;
  static_assert(R == 1 || R == 2 || R == 4 || R == 8, "rows per batch must be 1, 2, 4 or 8");
  const double k = RowsRed<R, W / 2, IS_MAX>::run(v, lane);
  if constexpr (TL::WPG > 1) {
    double *buf = scratch + phase * (TL::NW * R);
    phase ^= 1;
    if ((lane & (32 / R - 1)) == 0) buf[warp * R + row_of_lane<32, R>(lane)] = k;   // the canonical holder of every row
    group_sync<TL>(warp / TL::WPG);
    double tot = IS_MAX ? -INFINITY : 0.0;
    if (lane < R) {
      const int w0 = (warp / TL::WPG) * TL::WPG;
#pragma unroll
      for (int ww = 0; ww < TL::WPG; ++ww) tot = red_op<IS_MAX>(tot, buf[(w0 + ww) * R + lane]);
    }
    return tot;
  } else {
    return __shfl_sync(0xffffffffu, k, holder_lane<W, R>((lane % W) & (R - 1)), W);
  }
}

// =====================================================================================================
// Tails: the last CTA of a sweep sums the per-CTA partial vectors and, on one GPU, takes the control step.
// =====================================================================================================
// True in exactly one CTA of the launch: the one whose ticket shows that every other CTA has stored (and fenced)
// its partial vector.  Must be reached by all threads of every CTA that did not exit early.
__device__ __forceinline__ bool cta_is_last(ViCtl *ctl) {
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&ctl->ticket, 1u) == gridDim.x - 1 ? 1 : 0;
  __syncthreads();
  const bool last = s_last != 0;
  if (last) __threadfence();
  return last;
}

// Grid rendezvous of the fused cooperative kernels (every CTA resident): each CTA takes a ticket; the last one runs
// `last_does`, resets the ticket and bumps ViCtl.epoch; the others wait for the bump.  Fences on both sides as in cooperative
// groups' grid.sync(): what a CTA wrote before arriving is visible to every CTA after leaving, plain loads included.
template <class F>
__device__ __forceinline__ void grid_rendezvous(ViCtl *ctl, unsigned &epoch, F &&last_does) {
  if (cta_is_last(ctl)) {
    last_does();
    __syncthreads();
    if (threadIdx.x == 0) {
      ctl->ticket = 0;
      __threadfence();
      atomicExch(&ctl->epoch, epoch + 1u);
    }
  } else if (threadIdx.x == 0) {
    while (*reinterpret_cast<volatile unsigned *>(&ctl->epoch) != epoch + 1u) {}
    __threadfence();
  }
  __syncthreads();
  epoch += 1u;
}

// red[v] = sum over CTAs of partials[cta][v], v < nvals, CTAs in ascending order (whoever runs this: same bits).
__device__ __forceinline__ void reduce_partials(const double *partials, int pstride, int n_ctas, int nvals, double *red,
                                                int v_begin, int v_stride) {
  for (int v = v_begin; v < nvals; v += v_stride) {
    double a = 0.0;
    const double *p = partials + v;
#pragma unroll 8
    for (int c = 0; c < n_ctas; ++c) a += __ldcg(p + (size_t)c * pstride);
    red[v] = a;
  }
}
// The same sum taken by ONE CTA (the last of a sweep): a warp per column, its lanes stride over the CTAs (every load
// independent of the others: one round trip to L2 instead of a chain), then a shuffle sum — a fixed order for a
// given (NT, n_ctas), so run-to-run reproducible; it differs from the column-serial order of reduce_partials only
// in the last bits.
template <int NT>
__device__ __forceinline__ void reduce_partials_cta(const double *partials, int pstride, int n_ctas, int nvals, double *red,
                                                    double * /*seg: unused scratch*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int v = warp; v < nvals; v += NT / 32) {
    double a = 0.0;
    const double *p = partials + v;
#pragma unroll 4
    for (int c = lane; c < n_ctas; c += 32) a += __ldcg(p + (size_t)c * pstride);
    a = warp_sum(a);
    if (lane == 0) red[v] = a;
  }
}

// The same sum spread over a GRID (finalize kernels, the fused cooperative kernels: many CTAs x many columns).  A tile is 32
// columns; EIGHT warps share a tile: warp w adds the partial vectors w, w + 8, ... in ascending order — 256-byte coalesced
// rows, eight independent loads in flight per lane instead of one chain over all CTAs — and the eight sums are added in warp
// order.  A CTA of NT threads works on NT / 256 tiles at a time.  The order depends on n_ctas only (not on NT, not on the
// grid that runs it): run-to-run reproducible, and the same bits from the finalize kernels (1024 threads) and from the
// fused kernels (256 threads).  s_tile: NT doubles of shared memory.
template <int NT>
__device__ __forceinline__ void reduce_partials_tiled(const double *partials, int pstride, int n_ctas, int nvals, double *red,
                                                      double *s_tile) {
  static_assert(NT % 256 == 0, "eight warps per tile");
  constexpr int GROUPS = NT / 256;
  const int lane = threadIdx.x & 31, warp = (threadIdx.x >> 5) & 7, group = threadIdx.x >> 8;
  double *tile = s_tile + group * 256;
  for (int v0 = ((int)blockIdx.x * GROUPS) * 32; v0 < nvals; v0 += (int)gridDim.x * GROUPS * 32) {
    const int v = v0 + group * 32 + lane;
    double a = 0.0;
    if (v < nvals) {
      const double *p = partials + v;
#pragma unroll 8
      for (int c = warp; c < n_ctas; c += 8) a += __ldcg(p + (size_t)c * pstride);
    }
    __syncthreads();
    tile[warp * 32 + lane] = a;
    __syncthreads();
    if (warp == 0 && v < nvals) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += tile[w * 32 + lane];
      red[v] = t;
    }
  }
}
inline int finalize_grid(int nvals, int nt) { return (nvals + 32 * (nt / 256) - 1) / (32 * (nt / 256)); }

__device__ __forceinline__ void trace_push(const ViArrays &a, ViCtl *ctl, double gnorm, int reset) {
  if (ctl->iter < a.trace_cap) {
    a.trace_bound[ctl->iter] = ctl->bound;
    a.trace_gnorm[ctl->iter] = gnorm;
    a.trace_reset[ctl->iter] = (unsigned char)reset;
  }
}

// EM: N_k, bound, convergence test, then the next digamma / weight vector.  All NT threads of one CTA.
//   dense pass : red[k] = sum_j P(j,k) c_j / S_j (without w_k)                       N_k = alpha0 + w_k red[k]
//   sparse pass: red[k] = sum over the hits of k of their extra responsibilities (with w_k),
//                red[K + RED_AUX] = Z = sum_j P0_j c_j / S_j                          N_k = alpha0 + red[k] + w_k Z
//   red[K + RED_BOUND] = sum_j c_j (log S_j + M_j).
// (red[] is read with ld.cg: other CTAs of the same launch may have written it.)
template <int NT>
__device__ __forceinline__ void em_ctl_step(const ViArrays &a, ViCtl *ctl, int K, int sparse, double *scratch) {
  double lg = 0.0, dga = 0.0;
  const double z = __ldcg(a.red + K + RED_AUX);
  for (int k = threadIdx.x; k < K; k += NT) {
    const double rk = __ldcg(a.red + k);
    const double A = sparse ? rk + a.w[k] * z : a.w[k] * rk;
    const double nk = a.alpha0[k] + A;
    a.N_k[k] = nk;
    lg += lgamma(nk);
    dga += a.dg[k] * A;
  }
  lg = block_sum<NT>(lg, scratch);
  dga = block_sum<NT>(dga, scratch);
  double mx = -INFINITY;
  for (int k = threadIdx.x; k < K; k += NT) {
    const double dg = digamma_series(a.N_k[k]);
    a.dg_prev[k] = a.dg[k];
    a.dg[k] = dg;
    mx = fmax(mx, dg);
  }
  mx = block_max<NT>(mx, scratch);
  for (int k = threadIdx.x; k < K; k += NT) a.w[k] = exp(a.dg[k] - mx);
  if (threadIdx.x == 0) {
    // log-normaliser of class j is log S_j + M_j + dg_max, and sum_k q (logl - gamma) = lse_j - sum_k q dg_k
    const double data = __ldcg(a.red + K + RED_BOUND) + ctl->sum_counts * ctl->dg_max - dga;
    const double bound = data + lg + ctl->bound_const;
    ctl->oldbound = ctl->bound;
    ctl->bound = bound;
    trace_push(a, ctl, 0.0, 0);
    ctl->iter += 1;
    if (__ldcg(a.red + K + RED_FAULT) > 0.0) ctl->fault = 1;       // summed over ranks: the same decision everywhere
    if (ctl->iter > 1 && fabs(bound - ctl->oldbound) < ctl->tol) ctl->converged = 1;
    if (ctl->converged || ctl->iter >= ctl->max_iters || ctl->fault) ctl->done = 1;
    ctl->dg_max = mx;
  }
}

// RCG, after sweep B (stage 0) or after the restart sweep (stage 1).  All NT threads of one CTA.
// red[k] = sum_j c_j q(j,k), red[K + RED_BOUND] = sum_jk c_j q (logl - gamma), red[K + RED_AUX] = the gradient norm
// sweep A found (stage 0; the Fletcher-Reeves ratio sweep B used is recomputed from it and committed here).
// stall_on_reject: several GPUs — a rejected step pauses the iterations (ctl->stall) until the host has enqueued the
// restart; one GPU — the restart sweep sits behind every step and runs when ctl->didreset says so.
template <int NT>
__device__ __forceinline__ void rcg_ctl_b_step(const ViArrays &a, ViCtl *ctl, int K, int stage, int stall_on_reject,
                                               double *scratch) {
  __shared__ int s_accept;
  double lg = 0.0;
  for (int k = threadIdx.x; k < K; k += NT) lg += lgamma(a.alpha0[k] + __ldcg(a.red + k));
  lg = block_sum<NT>(lg, scratch);
  const double cand = __ldcg(a.red + K + RED_BOUND) + lg + ctl->bound_const;
  if (threadIdx.x == 0) {
    if (stage == 0) {
      const double nn = __ldcg(a.red + K + RED_AUX);
      ctl->beta = nn / ctl->oldnorm;
      ctl->newnorm = nn;
      ctl->oldnorm = nn;
      ctl->didreset = 0;        // the flag of the PREVIOUS iteration has been consumed by sweep B
    }
    if (stage == 0 && cand < ctl->bound) {
      // the conjugate direction lost ground: drop it, redo the step from the same N_k (restart sweep)
      ctl->didreset = 1;
      ctl->resets += 1;
      if (stall_on_reject) ctl->stall = 1;
      s_accept = 0;
    } else {
      s_accept = 1;
      ctl->oldbound = ctl->bound;
      ctl->bound = cand;
      trace_push(a, ctl, ctl->newnorm, stage);
      ctl->iter += 1;
      ctl->stall = 0;
      if (stage == 0 && cand - ctl->oldbound < ctl->tol) ctl->converged = 1;
      if (ctl->converged || ctl->iter >= ctl->max_iters) ctl->done = 1;
    }
  }
  __syncthreads();
  if (!s_accept) return;
  for (int k = threadIdx.x; k < K; k += NT) {
    const double nk = a.alpha0[k] + __ldcg(a.red + k);
    a.N_k[k] = nk;
    a.dg[k] = digamma_series(nk) - 1.0;
  }
}

// What an EM sweep does once its own partial vector is stored (sparse: 0 dense pass, 1 sparse pass).
template <int NT>
__device__ __forceinline__ void em_sweep_tail(int tail, int sparse, const ViArrays &a, ViCtl *ctl, const double *partials,
                                              int pstride, int K, double *scratch) {
  if (tail == 0) return;
  if (!cta_is_last(ctl)) return;
  reduce_partials_cta<NT>(partials, pstride, (int)gridDim.x, K + RED_EXTRA, a.red, a.seg);
  if (threadIdx.x == 0) ctl->ticket = 0;
  if (tail == 2) {
    __syncthreads();
    em_ctl_step<NT>(a, ctl, K, sparse, scratch);
  }
}

// =====================================================================================================
// EM / VB pass, linear domain.  P(j,k) = exp(logl(j,k) - M_j) is stored once; a pass is two GEMVs that
// share one read of P:   S_j = sum_k P(j,k) w_k ,  A_k = sum_j P(j,k) c_j / S_j ,  with
// w_k = exp(digamma(N_k) - max digamma).  Then N_k = alpha0_k + w_k A_k and the data term of the ELBO
// is sum_j c_j (log S_j + M_j) (+ constants applied by the control step).  Two FMAs per element.
//
// P, counts and rowmax are padded with zero rows to a multiple of ROW_PAD classes: a padded class
// has c = 0 and drops out, so the sweep carries no row-bounds predicates at all.
// =====================================================================================================

// Two CTAs per SM when the two register buffers, the accumulators and the weights leave room within 128
// registers per thread (estimate in 32-bit registers: 2 buffers x R x KITER x 4, KITER x VEC doubles, KITER x VEC weights).
template <typename ST, class TL> constexpr int em_min_blocks() {
  constexpr int VEC = 16 / (int)sizeof(ST);
  constexpr int est = 8 * TL::R * TL::KITER + 2 * TL::KITER * VEC + TL::KITER * VEC * ((int)sizeof(ST) / 4);
  if (TL::NT > 256) return 1;
  if (est <= 100) return 2;
  // 160 / 192-thread CTAs still fit twice (204 / 170 registers each) — except fp32 with eight pieces, which needs ~240
  return TL::NT <= 192 && (TL::KITER <= 4 || sizeof(ST) == 8) ? 2 : 1;
}

// TAIL: whether the last-CTA reduction / control step is compiled in.  Its code (lgamma, digamma, block reductions)
// perturbs the scheduling of the streaming loads of the hot loop enough to cost 12 % on matrices of tens of GB, so the
// large problems — which use finalize_ctl_kernel anyway — run the TAIL = false instantiation.
template <typename ST, class TL, bool PIPE, bool TAIL>
__global__ void __launch_bounds__(TL::NT, em_min_blocks<ST, TL>())
em_lin_pass_kernel(const ST *__restrict__ P, int ld, const double *__restrict__ rowmax, const double *__restrict__ counts,
                   ViArrays arrays, ViCtl *ctl, double *partials, int pstride, unsigned long long N_pad, int K,
                   PipeGeom geom, int tail) {
  using VT = typename VecOf<ST>::type;
  constexpr int VEC = VecOf<ST>::VEC;
  constexpr int R = TL::R, KITER = TL::KITER, TPR = TL::TPR, WPG = TL::WPG, W = TL::W;
  constexpr int RB = TL::G * R;                       // rows per CTA batch
  static_assert(ROW_PAD % RB == 0, "ROW_PAD must be a multiple of the rows of one CTA batch");
  if (ctl->done) return;
  __shared__ double s_red[2 * TL::NW * R];
  __shared__ double s_comb[TL::G > 1 ? TPR * KITER * VEC : 1];
  __shared__ double s_blk[32];
  const double *w = arrays.w;                         // (rewritten by the control step of the tail: no __restrict__)
  const int t = threadIdx.x % TPR, g = threadIdx.x / TPR;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tl = lane % W;                            // lane inside the segment that shares the rows
  const int nvec = ld / VEC;

  ST wv[KITER][VEC];
  bool inr[KITER];
#pragma unroll
  for (int i = 0; i < KITER; ++i) {
    inr[i] = t + TPR * i < nvec;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const int k = VEC * (t + TPR * i) + v;
      wv[i][v] = k < K ? (ST)w[k] : (ST)0;
    }
  }
  double acc[KITER][VEC];
#pragma unroll
  for (int i = 0; i < KITER; ++i)
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[i][v] = 0.0;
  double elbo = 0.0;
  int fault = 0;
  int phase = 0;

  // Which row of the batch this lane normalises.  Rows that span warps: lanes 0..R-1 of every warp (the first warp
  // of the group adds the bound term).  Rows inside a warp: every lane the row its transposing bits spell (W / R lanes
  // per row compute the same division; the canonical holder adds the bound term).
  const int my_row = WPG > 1 ? lane : row_of_lane<W, R>(lane);
  const bool my_active = WPG > 1 ? lane < R : true;
  const bool my_elbo = WPG > 1 ? t < 32 : tl == holder_lane<W, R>(my_row);

  // ---- one batch: pv = this thread's pieces of R rows, c_lane = count of row (row0 + my_row)
  auto process = [&](const VT (&pv)[R][KITER], double c_lane, unsigned long long row0) {
    double s[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      ST a[VEC];   // fp32 storage: the row-local dot product runs in fp32, everything across rows in fp64
#pragma unroll
      for (int v = 0; v < VEC; ++v) a[v] = (ST)0;
#pragma unroll
      for (int i = 0; i < KITER; ++i) {
        ST e[VEC];
        unpack(pv[r][i], e);
#pragma unroll
        for (int v = 0; v < VEC; ++v) a[v] = fma(e[v], wv[i][v], a[v]);
      }
      ST tot = a[0];
#pragma unroll
      for (int v = 1; v < VEC; ++v) tot += a[v];
      s[r] = (double)tot;
    }
    const double k = RowsRed<R, W / 2, false>::run(s, lane);
    double tot;
    if constexpr (WPG > 1) {
      double *buf = s_red + phase * (TL::NW * R);
      phase ^= 1;
      if ((lane & (32 / R - 1)) == 0) buf[warp * R + row_of_lane<32, R>(lane)] = k;
      group_sync<TL>(g);
      tot = 0.0;
      if (lane < R) {
        const int w0 = (warp / WPG) * WPG;
#pragma unroll
        for (int ww = 0; ww < WPG; ++ww) tot += buf[(w0 + ww) * R + lane];
      }
    } else {
      tot = k;                                  // every lane already holds the total of row my_row
    }
    // One division per row serves the whole batch.
    double inv_l = 0.0;
    if (my_active && c_lane > 0.0) {
      if (!(tot > 0.0) || isinf(tot)) fault = 1;
      else {
        inv_l = c_lane / tot;
        if (my_elbo) elbo += c_lane * (log(tot) + rowmax[row0 + my_row]);
      }
    }
    ST inv[R];
#pragma unroll
    for (int r = 0; r < R; ++r) inv[r] = (ST)__shfl_sync(0xffffffffu, inv_l, WPG > 1 ? r : holder_lane<W, R>(r), W);
    // A_k += sum_r P(r,k) c_r / S_r: the R-row partial in storage precision, the running sum in fp64
#pragma unroll
    for (int i = 0; i < KITER; ++i) {
      ST a[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v) a[v] = (ST)0;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        ST e[VEC];
        unpack(pv[r][i], e);
#pragma unroll
        for (int v = 0; v < VEC; ++v) a[v] = fma(e[v], inv[r], a[v]);
      }
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[i][v] += (double)a[v];
    }
  };

  const VT zero = VT{};
  if constexpr (!PIPE) {
    // direct 128-bit streaming loads, software-pipelined one batch ahead (two register buffers)
    const unsigned long long n_batches = N_pad / RB;
    const VT *base = reinterpret_cast<const VT *>(P) + t;
    auto fetch = [&](VT (&buf)[R][KITER], double &c_lane, unsigned long long b) {
      const unsigned long long row0 = b * RB + (unsigned long long)g * R;
      const VT *p = base + row0 * (unsigned long long)nvec;
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int i = 0; i < KITER; ++i)
          if (inr[i]) buf[r][i] = ld_stream(p + (size_t)r * nvec + i * TPR);
      c_lane = my_active ? counts[row0 + my_row] : 0.0;
    };
    VT bufA[R][KITER], bufB[R][KITER];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int i = 0; i < KITER; ++i) { bufA[r][i] = zero; bufB[r][i] = zero; }
    double cA = 0.0, cB = 0.0;
    // Batch -> CTA mapping.  Interleaved (stride = grid): the whole GPU reads one contiguous window at a time.
    // Chunked (geom.stage_rows == 1 in direct mode): every CTA walks its own contiguous range of batches, so an SM
    // stays inside one 2 MB page for many batches — far fewer TLB fills once the matrix outgrows the TLB reach.
    const bool chunked = geom.stage_rows == 1;
    const unsigned long long per = (n_batches + gridDim.x - 1) / gridDim.x;
    const unsigned long long stride = chunked ? 1 : gridDim.x;
    unsigned long long b = chunked ? blockIdx.x * per : blockIdx.x;
    const unsigned long long b_end = chunked ? min(n_batches, b + per) : n_batches;
    if (b < b_end) {
      const unsigned long long last = b_end - 1;
      fetch(bufA, cA, b);
      for (; b < b_end; b += 2 * stride) {
        const unsigned long long b1 = b + stride;
        fetch(bufB, cB, min(b1, last));                 // past the end: a harmless reload, never processed
        process(bufA, cA, b * RB + (unsigned long long)g * R);
        if (b1 < b_end) {
          fetch(bufA, cA, min(b1 + stride, last));
          process(bufB, cB, b1 * RB + (unsigned long long)g * R);
        }
      }
    }
  } else {
    RowPipe<1> pipe;
    pipe.buf = g_dyn_smem;
    pipe.full = reinterpret_cast<uint64_t *>(g_dyn_smem + (size_t)geom.stages * geom.stage_pitch);
    pipe.stages = geom.stages; pipe.stage_rows = geom.stage_rows; pipe.nsrc = 1;
    pipe.row_bytes = (uint32_t)ld * sizeof(ST); pipe.stage_pitch = geom.stage_pitch;
    pipe.src[0] = reinterpret_cast<const unsigned char *>(P);
    pipe.N = N_pad;
    pipe.run([&](int stage, unsigned long long stage_row0) {
      const unsigned char *sp = pipe.stage_ptr(stage, 0);
      const int rows_here = (int)min((unsigned long long)geom.stage_rows, N_pad - stage_row0);
      for (int lrow = 0; lrow < rows_here; lrow += RB) {
        const int lrow0 = lrow + g * R;
        VT pv[R][KITER];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int i = 0; i < KITER; ++i)
            pv[r][i] = inr[i] ? reinterpret_cast<const VT *>(sp + (size_t)(lrow0 + r) * pipe.row_bytes)[t + TPR * i] : zero;
        const unsigned long long row0 = stage_row0 + lrow0;
        process(pv, my_active ? counts[row0 + my_row] : 0.0, row0);
      }
    });
  }
  double *out = partials + (unsigned long long)blockIdx.x * pstride;
  store_partials<TL, VEC>(acc, s_comb, out, K);
  elbo = block_sum<TL::NT>(elbo, s_blk);
  const int any_fault = __syncthreads_or(fault);
  if (threadIdx.x == 0) { out[K + RED_BOUND] = elbo; out[K + RED_AUX] = 0.0; out[K + RED_FAULT] = any_fault ? 1.0 : 0.0; }
  if constexpr (TAIL) em_sweep_tail<TL::NT>(tail, 0, arrays, ctl, partials, pstride, K, s_blk);
}

// =====================================================================================================
// The same pass for a BATCH of bootstrap replicates (src/mSWEEP.cpp:496-518: the same likelihood, B resampled count
// vectors, B cold-start estimations).  The matrix is the expensive thing to read and it does not depend on the
// replicate: a CTA streams a row batch ONCE and serves BT replicates from it — S_jb = sum_k P_jk w_kb,
// A_kb += P_jk c_jb / S_jb — i.e. bytes per replicate-iteration fall BT-fold (fp64 FMA throughput then sets the price:
// 2 BT FMAs per element).  Slices of BT replicates sit on gridDim.y; CTAs of different slices walk the same rows at the
// same time, so the second slice's reads are served by L2.  Every replicate has its own weights, counts, control block
// and convergence; a replicate that has converged drops out of the `active` list at the host's next poll.
// Per-replicate vectors live at base + replicate * K (red: K + RED_EXTRA); traces are not kept.
// =====================================================================================================
__device__ __forceinline__ ViArrays arrays_of_replicate(const ViArrays &base, int rep, int K, int pstride) {
  ViArrays a = base;
  a.N_k += (size_t)rep * K; a.dg += (size_t)rep * K; a.w += (size_t)rep * K; a.dg_prev += (size_t)rep * K;
  a.red += (size_t)rep * (K + RED_EXTRA);
  a.seg += (size_t)rep * RED_SEGS * pstride;
  a.trace_cap = 0;
  return a;
}

template <typename ST, class TL, int BT>
__global__ void __launch_bounds__(TL::NT, (TL::NT <= 256 && 4 * TL::R * TL::KITER + 2 * BT * TL::KITER * (16 / (int)sizeof(ST)) + 2 * TL::R * BT <= 100) ? 2 : 1)
em_lin_batch_kernel(const ST *__restrict__ P, int ld, const double *__restrict__ rowmax, const double *__restrict__ counts,
                    unsigned long long counts_stride, ViArrays base, const int *__restrict__ active, int n_active,
                    double *partials, int pstride, unsigned long long N_pad, int K) {
  using VT = typename VecOf<ST>::type;
  constexpr int VEC = VecOf<ST>::VEC;
  constexpr int R = TL::R, KITER = TL::KITER, TPR = TL::TPR, WPG = TL::WPG;
  constexpr int NP = R * BT;                          // (row, replicate) pairs of one batch: one normaliser each
  constexpr int RB = TL::G * R;
  static_assert(TPR >= 32 && NP <= 32, "the batched pass uses whole-warp rows and at most 32 (row, replicate) pairs per batch");
  static_assert(ROW_PAD % RB == 0, "ROW_PAD must be a multiple of the rows of one CTA batch");
  ST *sW = reinterpret_cast<ST *>(g_dyn_smem);        // [BT][ld] weights of the slice's replicates
  __shared__ double s_red[2 * TL::NW * NP];
  __shared__ double s_comb[TL::G > 1 ? TPR * KITER * VEC : 1];
  __shared__ double s_blk[32];
  const int slice = blockIdx.y, b0 = slice * BT, nb = min(BT, n_active - b0);
  const int t = threadIdx.x % TPR, g = threadIdx.x / TPR;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nvec = ld / VEC;
  for (int idx = threadIdx.x; idx < BT * ld; idx += TL::NT) {
    const int b = idx / ld, k = idx - b * ld;
    sW[idx] = (b < nb && k < K) ? (ST)base.w[(size_t)active[b0 + b] * K + k] : (ST)0;
  }
  __syncthreads();

  bool inr[KITER];
#pragma unroll
  for (int i = 0; i < KITER; ++i) inr[i] = t + TPR * i < nvec;
  double acc[BT][KITER][VEC];
#pragma unroll
  for (int b = 0; b < BT; ++b)
#pragma unroll
    for (int i = 0; i < KITER; ++i)
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[b][i][v] = 0.0;
  double elbo = 0.0;
  int fault = 0, phase = 0;

  // the (row, replicate) pair this lane normalises: pair p = r * BT + b
  const int my_pair = WPG > 1 ? lane : row_of_lane<32, NP>(lane);
  const bool pair_on = WPG > 1 ? lane < NP : true;
  const int my_r = (my_pair % NP) / BT, my_b = (my_pair % NP) % BT;
  const bool my_elbo = WPG > 1 ? t < 32 : lane == holder_lane<32, NP>(my_pair);
  const double *my_counts = (pair_on && my_b < nb) ? counts + (size_t)active[b0 + my_b] * counts_stride : nullptr;

  const unsigned long long n_batches = N_pad / RB;
  const VT *pbase = reinterpret_cast<const VT *>(P) + t;
  const VT zero = VT{};
  for (unsigned long long bt = blockIdx.x; bt < n_batches; bt += gridDim.x) {
    const unsigned long long row0 = bt * RB + (unsigned long long)g * R;
    const VT *p = pbase + row0 * (unsigned long long)nvec;
    VT pv[R][KITER];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int i = 0; i < KITER; ++i) pv[r][i] = inr[i] ? ld_stream(p + (size_t)r * nvec + i * TPR) : zero;
    const double c_lane = my_counts ? my_counts[row0 + my_r] : 0.0;
    double s[NP];
#pragma unroll
    for (int b = 0; b < BT; ++b) {
      VT wv[KITER];
#pragma unroll
      for (int i = 0; i < KITER; ++i) wv[i] = inr[i] ? reinterpret_cast<const VT *>(sW + (size_t)b * ld)[t + TPR * i] : zero;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if constexpr (sizeof(ST) == 8) {
          // fp64: one FMA chain per (row, replicate), no temporaries
          double a = 0.0;
#pragma unroll
          for (int i = 0; i < KITER; ++i) { a = fma(pv[r][i].x, wv[i].x, a); a = fma(pv[r][i].y, wv[i].y, a); }
          s[r * BT + b] = a;
        } else {
          ST a[VEC];
#pragma unroll
          for (int v = 0; v < VEC; ++v) a[v] = (ST)0;
#pragma unroll
          for (int i = 0; i < KITER; ++i) {
            ST e[VEC], ww[VEC];
            unpack(pv[r][i], e);
            unpack(wv[i], ww);
#pragma unroll
            for (int v = 0; v < VEC; ++v) a[v] = fma(e[v], ww[v], a[v]);
          }
          ST tot = a[0];
#pragma unroll
          for (int v = 1; v < VEC; ++v) tot += a[v];
          s[r * BT + b] = (double)tot;
        }
      }
    }
    const double k = RowsRed<NP, 16, false>::run(s, lane);
    double tot;
    if constexpr (WPG > 1) {
      double *buf = s_red + phase * (TL::NW * NP);
      phase ^= 1;
      if ((lane & (32 / NP - 1)) == 0) buf[warp * NP + row_of_lane<32, NP>(lane)] = k;
      group_sync<TL>(g);
      tot = 0.0;
      if (lane < NP) {
        const int w0 = (warp / WPG) * WPG;
#pragma unroll
        for (int ww = 0; ww < WPG; ++ww) tot += buf[(w0 + ww) * NP + lane];
      }
    } else {
      tot = k;
    }
    double inv_l = 0.0;
    if (c_lane > 0.0) {
      if (!(tot > 0.0) || isinf(tot)) fault = 1;
      else {
        inv_l = c_lane / tot;
        if (my_elbo) elbo += c_lane * (log(tot) + rowmax[row0 + my_r]);
      }
    }
#pragma unroll
    for (int b = 0; b < BT; ++b) {
      ST inv[R];
#pragma unroll
      for (int r = 0; r < R; ++r) inv[r] = (ST)__shfl_sync(0xffffffffu, inv_l, WPG > 1 ? r * BT + b : holder_lane<32, NP>(r * BT + b));
#pragma unroll
      for (int i = 0; i < KITER; ++i) {
        if constexpr (sizeof(ST) == 8) {
#pragma unroll
          for (int r = 0; r < R; ++r) {
            acc[b][i][0] = fma(pv[r][i].x, inv[r], acc[b][i][0]);
            acc[b][i][1] = fma(pv[r][i].y, inv[r], acc[b][i][1]);
          }
        } else {
          ST a[VEC];
#pragma unroll
          for (int v = 0; v < VEC; ++v) a[v] = (ST)0;
#pragma unroll
          for (int r = 0; r < R; ++r) {
            ST e[VEC];
            unpack(pv[r][i], e);
#pragma unroll
            for (int v = 0; v < VEC; ++v) a[v] = fma(e[v], inv[r], a[v]);
          }
#pragma unroll
          for (int v = 0; v < VEC; ++v) acc[b][i][v] += (double)a[v];
        }
      }
    }
  }
  // one partial vector per (replicate slot, CTA): partials[(slot * gridDim.x + cta) * pstride + ...]
#pragma unroll
  for (int b = 0; b < BT; ++b) {
    double *out = partials + ((size_t)(b0 + b) * gridDim.x + blockIdx.x) * pstride;
    if (b < nb) store_partials<TL, VEC>(acc[b], s_comb, out, K);       // (nb is uniform over the CTA)
    const double eb = block_sum<TL::NT>((pair_on && my_b == b) ? elbo : 0.0, s_blk);
    const int fb = __syncthreads_or((fault && my_b == b) ? 1 : 0);
    if (b < nb && threadIdx.x == 0) { out[K + RED_BOUND] = eb; out[K + RED_AUX] = 0.0; out[K + RED_FAULT] = fb ? 1.0 : 0.0; }
  }
}

// Reduction + control step of the batched pass: gridDim.y = active replicates, the last CTA of a replicate's row
// of CTAs takes its control step.
static __global__ void __launch_bounds__(256)
finalize_ctl_batch_kernel(const double *partials, int pstride, int n_ctas, int nvals, ViArrays base, ViCtl *ctls, int K,
                          const int *__restrict__ active) {
  const int slot = blockIdx.y, rep = active[slot];
  ViCtl *ctl = ctls + rep;
  if (ctl->done) return;
  __shared__ double s_blk[32];
  __shared__ double s_tile[256];
  const ViArrays a = arrays_of_replicate(base, rep, K, pstride);
  reduce_partials_tiled<256>(partials + (size_t)slot * n_ctas * pstride, pstride, n_ctas, nvals, a.red, s_tile);
  if (!cta_is_last(ctl)) return;
  if (threadIdx.x == 0) ctl->ticket = 0;
  __syncthreads();
  em_ctl_step<256>(a, ctl, K, 0, s_blk);
}

// =====================================================================================================
// Log-domain sweeps (fp64): the RCG optimiser and its restart step.
// =====================================================================================================

// Sweep A of an RCG iteration ("mixt_negnatgrad"): d = logl + (digamma(N_k) - 1) - gamma,
// newnorm = sum_jk q (d - <d>_j) d  with q = exp(gamma), <d>_j = sum_k q d.  Nothing is written: d is
// recomputed by sweep B, which saves 16 B/element of traffic over storing it.  The last CTA to finish sums
// the per-CTA norms into red[K + RED_AUX] (one value per CTA: always cheap).
template <class TL, bool PIPE>
__global__ void __launch_bounds__(TL::NT, TL::NT > 256 ? 1 : (TL::R * TL::KITER <= 4 ? 3 : (TL::R * TL::KITER <= 8 ? 2 : 1)))
rcg_sweep_a_kernel(const double *__restrict__ logl, const double *__restrict__ gamma, int ld,
                   ViArrays arrays, ViCtl *ctl, double *partials,
                   int pstride, unsigned long long N, int K, PipeGeom geom) {
  constexpr int R = TL::R, KITER = TL::KITER, TPR = TL::TPR, W = TL::W;
  if (ctl->done || ctl->stall) return;
  __shared__ double s_red[2 * TL::NW * R];
  __shared__ double s_blk[32];
  const double *dgm1 = arrays.dg;
  const int t = threadIdx.x % TPR, g = threadIdx.x / TPR;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nvec = ld / 2;
  double dk[KITER][2];
  bool ok[KITER][2];
  bool cols_ok = true;
#pragma unroll
  for (int i = 0; i < KITER; ++i)
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      const int k = 2 * (t + TPR * i) + v;
      ok[i][v] = k < K;
      cols_ok = cols_ok && ok[i][v];
      dk[i][v] = ok[i][v] ? dgm1[k] : 0.0;
    }
  double nn = 0.0;
  int phase = 0;
  const bool warp_cols_ok = __all_sync(0xffffffffu, cols_ok) != 0;

  auto body = [&](auto &&load, unsigned long long row0) {   // load(src, r, idx): src 0 = logl, 1 = gamma
    double d[R][KITER][2], q[R][KITER][2];
    double s[R];
    // Three flavours (no shuffles inside: the row groups of a warp may take different ones at the ragged end):
    // every row and column real — no predicates; rows real, columns ragged (K not a multiple of the tile width) —
    // per-thread column flags only; ragged rows — everything predicated.
    if (row0 + R <= N) {
      if (warp_cols_ok) {
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int i = 0; i < KITER; ++i) { unpack(load(0, r, t + TPR * i), d[r][i]); unpack(load(1, r, t + TPR * i), q[r][i]); }
#pragma unroll
        for (int r = 0; r < R; ++r) {
          s[r] = 0.0;
#pragma unroll
          for (int i = 0; i < KITER; ++i)
#pragma unroll
            for (int v = 0; v < 2; ++v) {
              const double gam = q[r][i][v];
              d[r][i][v] = d[r][i][v] + dk[i][v] - gam;
              q[r][i][v] = exp_nonpos(gam);
              s[r] = fma(d[r][i][v], q[r][i][v], s[r]);
            }
        }
      } else {
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int i = 0; i < KITER; ++i) {
            if (t + TPR * i < nvec) { unpack(load(0, r, t + TPR * i), d[r][i]); unpack(load(1, r, t + TPR * i), q[r][i]); }
            else { d[r][i][0] = d[r][i][1] = 0.0; q[r][i][0] = q[r][i][1] = 0.0; }
          }
#pragma unroll
        for (int r = 0; r < R; ++r) {
          s[r] = 0.0;
#pragma unroll
          for (int i = 0; i < KITER; ++i)
#pragma unroll
            for (int v = 0; v < 2; ++v) {
              const double gam = q[r][i][v];
              d[r][i][v] = ok[i][v] ? d[r][i][v] + dk[i][v] - gam : 0.0;
              q[r][i][v] = ok[i][v] ? exp_nonpos(gam) : 0.0;
              s[r] = fma(d[r][i][v], q[r][i][v], s[r]);
            }
        }
      }
    } else {
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int i = 0; i < KITER; ++i) {
          const int idx = t + TPR * i;
          if (row0 + r < N && idx < nvec) { unpack(load(0, r, idx), d[r][i]); unpack(load(1, r, idx), q[r][i]); }
          else { d[r][i][0] = d[r][i][1] = 0.0; q[r][i][0] = q[r][i][1] = 0.0; }
        }
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
            d[r][i][v] = on ? d[r][i][v] + dk[i][v] - gam : 0.0;
            q[r][i][v] = on ? exp_nonpos(fmin(gam, 0.0)) : 0.0;
            s[r] = fma(d[r][i][v], q[r][i][v], s[r]);
          }
      }
    }
    const double tot = rows_reduce<TL, false>(s, s_red, phase, lane, warp);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const double sr = __shfl_sync(0xffffffffu, tot, r, W);
#pragma unroll
      for (int i = 0; i < KITER; ++i)
#pragma unroll
        for (int v = 0; v < 2; ++v) nn = fma(q[r][i][v] * (d[r][i][v] - sr), d[r][i][v], nn);
    }
  };

  if constexpr (!PIPE) {
    const unsigned long long rows_per_batch = (unsigned long long)TL::G * R;
    const unsigned long long n_batches = (N + rows_per_batch - 1) / rows_per_batch;
    for (unsigned long long b = blockIdx.x; b < n_batches; b += gridDim.x) {
      const unsigned long long row0 = b * rows_per_batch + (unsigned long long)g * R;
      const double2 *lp = reinterpret_cast<const double2 *>(logl) + row0 * (unsigned long long)nvec;
      const double2 *gp = reinterpret_cast<const double2 *>(gamma) + row0 * (unsigned long long)nvec;
      body([&](int s, int r, int idx) { return ld_stream((s == 0 ? lp : gp) + (size_t)r * nvec + idx); }, row0);
    }
  } else {
    RowPipe<2> pipe;
    pipe.buf = g_dyn_smem;
    pipe.full = reinterpret_cast<uint64_t *>(g_dyn_smem + (size_t)geom.stages * 2 * geom.stage_pitch);
    pipe.stages = geom.stages; pipe.stage_rows = geom.stage_rows; pipe.nsrc = 2;
    pipe.row_bytes = (uint32_t)ld * sizeof(double); pipe.stage_pitch = geom.stage_pitch;
    pipe.src[0] = reinterpret_cast<const unsigned char *>(logl);
    pipe.src[1] = reinterpret_cast<const unsigned char *>(gamma);
    pipe.N = N;
    const int sub = geom.stage_rows / (TL::G * R);
    pipe.run([&](int stage, unsigned long long stage_row0) {
      const unsigned char *sp0 = pipe.stage_ptr(stage, 0), *sp1 = pipe.stage_ptr(stage, 1);
      for (int sb = 0; sb < sub; ++sb) {
        const int lrow0 = (sb * TL::G + g) * R;
        body([&](int s, int r, int idx) {
          return reinterpret_cast<const double2 *>((s == 0 ? sp0 : sp1) + (size_t)(lrow0 + r) * pipe.row_bytes)[idx];
        }, stage_row0 + lrow0);
      }
    });
  }
  nn = block_sum<TL::NT>(nn, s_blk);
  if (threadIdx.x == 0) partials[(unsigned long long)blockIdx.x * pstride + K + RED_AUX] = nn;
  if (cta_is_last(ctl)) {
    double a = 0.0;
    for (int c = threadIdx.x; c < (int)gridDim.x; c += TL::NT) a += __ldcg(partials + (size_t)c * pstride + K + RED_AUX);
    a = block_sum<TL::NT>(a, s_blk);
    if (threadIdx.x == 0) { arrays.red[K + RED_AUX] = a; ctl->ticket = 0; }
  }
}

// Sweep B of an RCG iteration: step = d (+ beta * oldstep), gamma += step, renormalise every class,
// store gamma and step, accumulate N_k - alpha0 = sum_j c_j q and the data term of the ELBO.
// MODE 0: RCG step; the Fletcher-Reeves ratio beta = newnorm / oldnorm comes from the (all-reduced) norm of sweep A
// in red[K + RED_AUX], recomputed by every CTA (the control step after the sweep commits it).
// MODE 1: plain step from the current digamma vector, gamma = normalise(logl + dg) (the RCG restart).
template <class TL, int MODE, bool WRITE, bool PIPE, bool TAIL>
__global__ void __launch_bounds__(TL::NT, TL::R * TL::KITER <= 4 && TL::NT <= 256 ? 2 : 1)
rcg_sweep_b_kernel(const double *__restrict__ logl, double *__restrict__ gamma, double *__restrict__ step, int ld,
                   ViArrays arrays, const double *__restrict__ counts, ViCtl *ctl,
                   double *partials, int pstride, unsigned long long N, int K, int only_if_reset,
                   PipeGeom geom, int tail) {
  constexpr int R = TL::R, KITER = TL::KITER, TPR = TL::TPR, W = TL::W;
  if (ctl->done) return;
  if (only_if_reset ? !ctl->didreset : ctl->stall != 0) return;
  __shared__ double s_red[2 * TL::NW * R];
  __shared__ double s_comb[TL::G > 1 ? TPR * KITER * 2 : 1];
  __shared__ double s_blk[32];
  const double *dgv = arrays.dg;
  const int t = threadIdx.x % TPR, g = threadIdx.x / TPR;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tl = lane % W;
  // the direction memory is empty before the first accepted step (the reference starts it at zero) and after a restart
  const double beta = MODE == 0 ? arrays.red[K + RED_AUX] / ctl->oldnorm : 0.0;
  const bool use_old = MODE == 0 && !ctl->didreset && beta > 0.0 && ctl->iter > 0;
  const double beta_eff = use_old ? beta : 0.0;   // step = d + beta * oldstep; no direction memory -> beta 0
  const int nvec = ld / 2;
  const double NEG_INF = -INFINITY;
  double dk[KITER][2];
  bool ok[KITER][2];
  bool cols_ok = true;
#pragma unroll
  for (int i = 0; i < KITER; ++i)
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      const int k = 2 * (t + TPR * i) + v;
      ok[i][v] = k < K;
      cols_ok = cols_ok && ok[i][v];
      dk[i][v] = ok[i][v] ? dgv[k] : 0.0;
    }
  double acc[KITER][2];
#pragma unroll
  for (int i = 0; i < KITER; ++i) acc[i][0] = acc[i][1] = 0.0;
  double bound = 0.0;
  int phase = 0;

  // One batch = three element-wise phases separated by two row reductions.  The phases exist in three
  // compile-time flavours: every row and column this thread touches is real (no predicates, the loads of a batch issue
  // back to back); rows real but columns ragged (K not a multiple of the tile width: per-thread column flags only); and
  // the fully predicated one for the last rows.  The reductions (shuffles with a full mask, barriers) sit outside the
  // flavoured code, in common control flow.
  const bool warp_cols_ok = __all_sync(0xffffffffu, cols_ok) != 0;
  auto body = [&](auto &&load, unsigned long long row0) {   // load(src, r, idx): 0 = logl, 1 = gamma, 2 = old step
    const bool rows_ok = row0 + R <= N;
    double l[R][KITER][2], gn[R][KITER][2], m[R];
    double e[R][KITER][2], sum[R], lsum[R], cinv[R];

    // ROWS / COLS are literals at the call sites; the lambdas are inlined and specialised for them
    auto phase1 = [&](const bool ROWS, const bool COLS) {
      double st[R][KITER][2];
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int i = 0; i < KITER; ++i) {
          const int idx = t + TPR * i;
          const bool in = (ROWS || row0 + r < N) && (COLS || idx < nvec);
          if (in) unpack(load(0, r, idx), l[r][i]); else l[r][i][0] = l[r][i][1] = 0.0;
          if (MODE == 0) { if (in) unpack(load(1, r, idx), gn[r][i]); else gn[r][i][0] = gn[r][i][1] = 0.0; }
        }
      if (MODE == 0 && use_old) {
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int i = 0; i < KITER; ++i) {
            const int idx = t + TPR * i;
            const bool in = (ROWS || row0 + r < N) && (COLS || idx < nvec);
            if (in) unpack(load(2, r, idx), st[r][i]); else st[r][i][0] = st[r][i][1] = 0.0;
          }
      } else {
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int i = 0; i < KITER; ++i) st[r][i][0] = st[r][i][1] = 0.0;
      }
      // the new direction leaves for HBM straight away so that its registers die before the reductions
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const unsigned long long row = row0 + r;
        double2 *sp = reinterpret_cast<double2 *>(step) + row * (unsigned long long)nvec;
        m[r] = NEG_INF;
#pragma unroll
        for (int i = 0; i < KITER; ++i) {
          const int idx = t + TPR * i;
#pragma unroll
          for (int v = 0; v < 2; ++v) {
            const bool on = (ROWS || row < N) && (COLS || ok[i][v]);
            double g2;
            if (MODE == 0) {
              const double d = l[r][i][v] + dk[i][v] - gn[r][i][v];
              const double s = fma(beta_eff, st[r][i][v], d);
              st[r][i][v] = on ? s : 0.0;
              g2 = gn[r][i][v] + s;
            } else {
              g2 = l[r][i][v] + dk[i][v];
            }
            gn[r][i][v] = on ? g2 : NEG_INF;
            m[r] = fmax(m[r], gn[r][i][v]);
          }
          if (MODE == 0 && (ROWS || row < N) && (COLS || idx < nvec)) st_stream(sp + idx, make_double2(st[r][i][0], st[r][i][1]));
        }
      }
    };
    auto phase3 = [&](const bool ROWS, const bool COLS) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const unsigned long long row = row0 + r;
        if (!ROWS && row >= N) continue;
        double2 *gp = reinterpret_cast<double2 *>(gamma) + row * (unsigned long long)nvec;
#pragma unroll
        for (int i = 0; i < KITER; ++i) {
          const int idx = t + TPR * i;
          if (!COLS && idx >= nvec) continue;
          double gout[2];
#pragma unroll
          for (int v = 0; v < 2; ++v) {
            const double g2 = gn[r][i][v] - lsum[r];     // normalised log-responsibility
            gout[v] = (COLS || ok[i][v]) ? g2 : 0.0;
            const double cq = cinv[r] * e[r][i][v];      // c_j q(j,k); 0 for unobserved classes and padding
            acc[i][v] += cq;
            if (cq > 0.0) bound = fma(cq, l[r][i][v] - g2, bound);
          }
          if (WRITE) st_stream(gp + idx, make_double2(gout[0], gout[1]));
        }
      }
    };

    if (rows_ok) { if (warp_cols_ok) phase1(true, true); else phase1(true, false); } else phase1(false, false);
    const double mt = rows_reduce<TL, true>(m, s_red, phase, lane, warp);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      double mm = __shfl_sync(0xffffffffu, mt, r, W);
      mm = mm == NEG_INF ? 0.0 : mm;                     // a row past the end: keep the arithmetic finite
      sum[r] = 0.0;
#pragma unroll
      for (int i = 0; i < KITER; ++i)
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          gn[r][i][v] -= mm;
          e[r][i][v] = exp_nonpos(gn[r][i][v]);          // 0 on padding columns (-inf)
          sum[r] += e[r][i][v];
        }
    }
    const double tot = rows_reduce<TL, false>(sum, s_red, phase, lane, warp);
    // lanes 0..R-1 of the segment: one log, one division and one count load per row serve the whole batch
    double lsum_l = 0.0, cinv_l = 0.0;
    if (tl < R && row0 + tl < N) {
      lsum_l = log(tot);
      cinv_l = counts[row0 + tl] / tot;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      lsum[r] = __shfl_sync(0xffffffffu, lsum_l, r, W);
      cinv[r] = __shfl_sync(0xffffffffu, cinv_l, r, W);
    }
    if (rows_ok) { if (warp_cols_ok) phase3(true, true); else phase3(true, false); } else phase3(false, false);
  };

  if constexpr (!PIPE) {
    const unsigned long long rows_per_batch = (unsigned long long)TL::G * R;
    const unsigned long long n_batches = (N + rows_per_batch - 1) / rows_per_batch;
    for (unsigned long long b = blockIdx.x; b < n_batches; b += gridDim.x) {
      const unsigned long long row0 = b * rows_per_batch + (unsigned long long)g * R;
      const unsigned long long off = row0 * (unsigned long long)nvec;
      const double2 *lp = reinterpret_cast<const double2 *>(logl) + off;
      const double2 *gp = reinterpret_cast<const double2 *>(gamma) + off;
      const double2 *sp = reinterpret_cast<const double2 *>(step) + off;
      body([&](int s, int r, int idx) { return ld_stream((s == 0 ? lp : (s == 1 ? gp : sp)) + (size_t)r * nvec + idx); }, row0);
    }
  } else {
    RowPipe<3> pipe;
    pipe.buf = g_dyn_smem;
    pipe.full = reinterpret_cast<uint64_t *>(g_dyn_smem + (size_t)geom.stages * 3 * geom.stage_pitch);
    pipe.stages = geom.stages; pipe.stage_rows = geom.stage_rows;
    pipe.nsrc = MODE == 0 ? (use_old ? 3 : 2) : 1;
    pipe.row_bytes = (uint32_t)ld * sizeof(double); pipe.stage_pitch = geom.stage_pitch;
    pipe.src[0] = reinterpret_cast<const unsigned char *>(logl);
    pipe.src[1] = reinterpret_cast<const unsigned char *>(gamma);
    pipe.src[2] = reinterpret_cast<const unsigned char *>(step);
    pipe.N = N;
    const int sub = geom.stage_rows / (TL::G * R);
    pipe.run([&](int stage, unsigned long long stage_row0) {
      const unsigned char *sp0 = pipe.stage_ptr(stage, 0), *sp1 = pipe.stage_ptr(stage, 1), *sp2 = pipe.stage_ptr(stage, 2);
      for (int sb = 0; sb < sub; ++sb) {
        const int lrow0 = (sb * TL::G + g) * R;
        body([&](int s, int r, int idx) {
          const unsigned char *base = s == 0 ? sp0 : (s == 1 ? sp1 : sp2);
          return reinterpret_cast<const double2 *>(base + (size_t)(lrow0 + r) * pipe.row_bytes)[idx];
        }, stage_row0 + lrow0);
      }
    });
  }
  double *out = partials + (unsigned long long)blockIdx.x * pstride;
  store_partials<TL, 2>(acc, s_comb, out, K);
  bound = block_sum<TL::NT>(bound, s_blk);
  if (threadIdx.x == 0) out[K + RED_BOUND] = bound;
  // Only the K per-group sums and the bound term are reduced here: slot RED_AUX of red[] holds the norm of sweep A,
  // which the control step still needs.
  if constexpr (TAIL) {
    if (tail != 0 && cta_is_last(ctl)) {
      reduce_partials_cta<TL::NT>(partials, pstride, (int)gridDim.x, K + 1, arrays.red, arrays.seg);
      if (threadIdx.x == 0) ctl->ticket = 0;
      if (tail == 2) {
        __syncthreads();
        rcg_ctl_b_step<TL::NT>(arrays, ctl, K, MODE == 0 ? 0 : 1, 0, s_blk);
      }
    }
  }
}

// =====================================================================================================
// EM / VB pass over the SPARSE storage: class j is P0_j on every group except its hits, where it is P0_j + dP.
//   S_j = P0_j W + sum_hits dP w_g                 (W = sum_k w_k)
//   N_k - alpha0_k = w_k Z + sum_{hits of k} q     (Z = sum_j P0_j c_j / S_j, the part every group receives;
//                                                   q = dP w_g c_j / S_j, the hit's extra responsibility)
// One WARP per chunk of 32 consecutive classes.  The chunk's hits are one contiguous range of the hit arrays:
// they are read with coalesced loads (hit-parallel), multiplied by the weight of their group and parked in shared
// memory; lane l then adds up the hits of class l in list order (class-parallel) and takes the division and the
// logarithm of its class; the scatter into the per-group accumulators is hit-parallel again.
// The accumulators are 64-bit FIXED-POINT sums in shared memory, updated with two native 32-bit shared atomics
// (low word, then high word + carry): |q| <= c_j, so with the scale 2^61 / 2^ceil(log2(sum_j c_j)) nothing can
// overflow, integer addition is associative — the pass is bit-reproducible whatever order the atomics land in —
// and no compare-and-swap loop is involved (sm_100 has no native 64-bit shared-memory add, integer or floating).
// =====================================================================================================
constexpr int SP_NT = 256;            // threads per CTA
constexpr int SP_CHUNK = 32;          // classes per warp chunk
constexpr int SP_STAGE = 224;         // hits a warp parks in shared memory at a time (a chunk averages ~4 per class)
constexpr int SP_PF = 4;              // slabs of 32 hits of the NEXT chunk kept in flight in registers
constexpr uint32_t SP_GRP_MASK = 0x00ffffffu;   // nz_grp: group in the low 24 bits, (class index & 31) in the top 8

__device__ __forceinline__ void fx_atomic_add(unsigned *acc2 /* [lo, hi] */, long long x) {
  const unsigned lo = (unsigned)(unsigned long long)x, hi = (unsigned)((unsigned long long)x >> 32);
  const unsigned old = atomicAdd(&acc2[0], lo);
  const unsigned carry = (old + lo) < old ? 1u : 0u;       // unsigned wrap-around of the low word
  if (hi + carry != 0u) atomicAdd(&acc2[1], hi + carry);
}

// A warp walks a CONTIGUOUS run of chunks, so the hits of its next chunk start where those of the current one end:
// while the arithmetic of chunk i runs, the per-class values of chunk i + 1 and the first SP_PF x 32 hits behind the
// current range are already in flight (the pass is bound by load latency, not by instructions).
// (the body is shared with ems_fused_kernel; it ends with the CTA's partial vector stored)
__device__ __forceinline__ void
em_sparse_pass_body(const uint64_t *__restrict__ nz_ptr, const uint32_t *__restrict__ nz_grp, const double *__restrict__ nz_dP,
                    const double *__restrict__ P0, const double *__restrict__ cm_const, const double *__restrict__ counts,
                    const ViArrays &arrays, double *partials, int pstride,
                    unsigned long long N, unsigned long long nnz, int K, double fx_scale, double *s_blk) {
  extern __shared__ __align__(16) unsigned char s_dyn_sp[];
  double *s_w = reinterpret_cast<double *>(s_dyn_sp);                         // [K]
  unsigned *s_acc = reinterpret_cast<unsigned *>(s_w + K);                    // [K][2] fixed-point accumulators
  double *s_val = reinterpret_cast<double *>(s_acc + 2 * (size_t)K);          // [warps][SP_STAGE] dP * w of the parked hits
  double *s_r = s_val + (SP_NT / 32) * SP_STAGE;                              // [warps][32] c_j / S_j of the chunk
  uint32_t *s_key = reinterpret_cast<uint32_t *>(s_r + (SP_NT / 32) * 32);    // [warps][SP_STAGE] packed (class, group)
  const double *w = arrays.w;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double wsum = 0.0;
  for (int k = threadIdx.x; k < K; k += SP_NT) { const double x = w[k]; s_w[k] = x; s_acc[2 * k] = 0u; s_acc[2 * k + 1] = 0u; wsum += x; }
  wsum = block_sum<SP_NT>(wsum, s_blk);        // same order in every CTA and on every rank
  __syncthreads();
  double *val = s_val + warp * SP_STAGE;
  uint32_t *key = s_key + warp * SP_STAGE;
  double *rr = s_r + warp * 32;
  double z = 0.0, elbo = 0.0;
  int fault = 0;
  const unsigned long long n_chunks = (N + SP_CHUNK - 1) / SP_CHUNK;
  const unsigned long long warps_total = (unsigned long long)gridDim.x * (SP_NT / 32);
  const unsigned long long per = (n_chunks + warps_total - 1) / warps_total;
  const unsigned long long ch_begin = min(n_chunks, ((unsigned long long)blockIdx.x * (SP_NT / 32) + warp) * per);
  const unsigned long long ch_end = min(n_chunks, ch_begin + per);
  const unsigned long long last_hit = nnz ? nnz - 1 : 0;

  struct ClassData { unsigned long long a, b; double c, p0; };
  auto load_class = [&](unsigned long long ch) {
    // everything the classes of a chunk need, requested together (coalesced: consecutive lanes, consecutive classes)
    const unsigned long long j = ch * SP_CHUNK + lane;
    const bool have = j < N;
    ClassData d;
    d.a = nz_ptr[have ? j : N]; d.b = nz_ptr[have ? j + 1 : N];
    d.c = have ? counts[j] : 0.0; d.p0 = have ? P0[j] : 0.0;
    return d;
  };
  uint32_t pg[SP_PF];
  double pd[SP_PF];
  auto load_hits = [&](unsigned long long base) {          // (clamped: reads past the end of the arrays never happen)
#pragma unroll
    for (int k = 0; k < SP_PF; ++k) {
      const unsigned long long e = min(base + (unsigned long long)(32 * k + lane), last_hit);
      pg[k] = nz_grp[e];
      pd[k] = nz_dP[e];
    }
  };
  auto normalise = [&](const ClassData &d, double s) {     // lane l: class l of the chunk
    double r = 0.0;
    if (d.c > 0.0) {
      if (!(s > 0.0) || isinf(s)) fault = 1;
      else {
        r = d.c / s;
        z = fma(r, d.p0, z);
        elbo = fma(d.c, log(s), elbo);          // (+ c_j M_j: the same in every pass — cm_const, added once below)
      }
    }
    __syncwarp();
    rr[lane] = r;
    __syncwarp();
  };
  auto scatter = [&](int n_here) {                          // hit-parallel
    for (int x = lane; x < n_here; x += 32) {
      const uint32_t kg = key[x];
      const double q = rr[kg >> 24] * val[x];
      if (q != 0.0) fx_atomic_add(&s_acc[2 * (kg & SP_GRP_MASK)], __double2ll_rn(q * fx_scale));
    }
  };

  if (ch_begin < ch_end) {
    ClassData nxt = load_class(ch_begin);
    load_hits(__shfl_sync(0xffffffffu, nxt.a, 0));           // (the one exposed round trip of the warp)
    for (unsigned long long ch = ch_begin; ch < ch_end; ++ch) {
      const ClassData cur = nxt;
      const unsigned long long h0 = __shfl_sync(0xffffffffu, cur.a, 0), h1 = __shfl_sync(0xffffffffu, cur.b, 31);
      if (ch + 1 < ch_end) nxt = load_class(ch + 1);
      double s = cur.p0 * wsum;
      if (h1 - h0 <= SP_STAGE) {
        const int n_here = (int)(h1 - h0);
        __syncwarp();
        // park the chunk's hits: the first SP_PF slabs arrived in registers, the rest comes straight from memory
        // (keeping key and value in registers for the scatter was measured: no gain — the extra live registers spill)
#pragma unroll
        for (int k = 0; k < SP_PF; ++k) {
          const int x = 32 * k + lane;
          if (x < n_here) { key[x] = pg[k]; val[x] = pd[k] * s_w[pg[k] & SP_GRP_MASK]; }
        }
        for (int x = 32 * SP_PF + lane; x < n_here; x += 32) {
          const uint32_t kg = nz_grp[h0 + x];
          key[x] = kg;
          val[x] = nz_dP[h0 + x] * s_w[kg & SP_GRP_MASK];
        }
        load_hits(h1);
        __syncwarp();
        // class-parallel: lane l adds the parked hits of class l in list order
        {
          const int xa = (int)(cur.a - h0), xe = (int)(cur.b - h0);
#pragma unroll
          for (int u = 0; u < 4; ++u) if (xa + u < xe) s += val[xa + u];      // (a class has a handful of hits: no loop for most lanes)
          for (int x = xa + 4; x < xe; ++x) s += val[x];
        }
        normalise(cur, s);
        scatter(n_here);
      } else {
        // a chunk with more hits than a stage holds: pieces, two rounds (the normalisers first, then the scatter)
        for (int round = 0; round < 2; ++round) {
          for (unsigned long long q0 = h0; q0 < h1; q0 += SP_STAGE) {
            const int n_here = (int)min((unsigned long long)SP_STAGE, h1 - q0);
            __syncwarp();
            for (int x = lane; x < n_here; x += 32) {
              const uint32_t kg = nz_grp[q0 + x];
              key[x] = kg;
              val[x] = nz_dP[q0 + x] * s_w[kg & SP_GRP_MASK];
            }
            __syncwarp();
            if (round == 0) {
              const unsigned long long lo = max(cur.a, q0), hi = min(cur.b, q0 + (unsigned long long)n_here);
              for (unsigned long long x = lo; x < hi; ++x) s += val[x - q0];
            } else {
              scatter(n_here);
            }
          }
          if (round == 0) normalise(cur, s);
        }
        load_hits(h1);
      }
    }
  }
  __syncthreads();
  double *out = partials + (unsigned long long)blockIdx.x * pstride;
  const double fx_inv = 1.0 / fx_scale;
  for (int k = threadIdx.x; k < K; k += SP_NT) {
    const long long v = (long long)(((unsigned long long)s_acc[2 * k + 1] << 32) | (unsigned long long)s_acc[2 * k]);
    out[k] = (double)v * fx_inv;
  }
  z = block_sum<SP_NT>(z, s_blk);
  elbo = block_sum<SP_NT>(elbo, s_blk);
  const int any_fault = __syncthreads_or(fault);
  // cm_const[0] = sum_j c_j M_j over this rank's classes (vi.cu: once per run; M_j = the row maximum the stored values are relative to)
  if (threadIdx.x == 0) { out[K + RED_BOUND] = blockIdx.x == 0 ? elbo + cm_const[0] : elbo; out[K + RED_AUX] = z; out[K + RED_FAULT] = any_fault ? 1.0 : 0.0; }
}
template <bool TAIL>
__global__ void __launch_bounds__(SP_NT, 3)
em_sparse_pass_kernel(const uint64_t *__restrict__ nz_ptr, const uint32_t *__restrict__ nz_grp, const double *__restrict__ nz_dP,
                      const double *__restrict__ P0, const double *__restrict__ cm_const, const double *__restrict__ counts,
                      ViArrays arrays, ViCtl *ctl, double *partials, int pstride,
                      unsigned long long N, unsigned long long nnz, int K, double fx_scale, int tail) {
  if (ctl->done) return;
  __shared__ double s_blk[32];
  em_sparse_pass_body(nz_ptr, nz_grp, nz_dP, P0, cm_const, counts, arrays, partials, pstride, N, nnz, K, fx_scale, s_blk);
  if constexpr (TAIL) em_sweep_tail<SP_NT>(tail, 1, arrays, ctl, partials, pstride, K, s_blk);
}
// Up to n_steps EM / VB iterations on the sparse storage in ONE cooperative launch (every CTA resident; one GPU): pass, grid
// rendezvous, reduction of the partial vectors (by the last arrival, or — coop_reduce — by every CTA on its tiles of columns
// followed by a second rendezvous), control step by the last arrival.  While a pass takes tens of microseconds (config 2 /
// 5: 33 us) the launches, their drains and a separate reduction kernel are a third of the iteration.  Same passes and the same
// order over the partial vectors as the launch-per-pass path; the control step's K-sized sums run over SP_NT instead of
// FIN_NT threads, so the two paths agree to the last digits of the bound (tests/test_gpu_scale.py).
static __global__ void __launch_bounds__(SP_NT, 3)
ems_fused_kernel(const uint64_t *__restrict__ nz_ptr, const uint32_t *__restrict__ nz_grp, const double *__restrict__ nz_dP,
                 const double *__restrict__ P0, const double *__restrict__ cm_const, const double *__restrict__ counts,
                 ViArrays arrays, ViCtl *ctl, double *partials, int pstride,
                 unsigned long long N, unsigned long long nnz, int K, double fx_scale, unsigned long long n_steps, int coop_reduce) {
  __shared__ double s_blk[32];
  __shared__ double s_tile[SP_NT];
  unsigned epoch = *reinterpret_cast<volatile unsigned *>(&ctl->epoch);
  for (unsigned long long it = 0; it < n_steps; ++it) {
    if (*reinterpret_cast<volatile int *>(&ctl->done)) break;          // (the same value in every CTA: read between rendezvous)
    em_sparse_pass_body(nz_ptr, nz_grp, nz_dP, P0, cm_const, counts, arrays, partials, pstride, N, nnz, K, fx_scale, s_blk);
    if (coop_reduce) {
      grid_rendezvous(ctl, epoch, [] {});
      reduce_partials_tiled<SP_NT>(partials, pstride, (int)gridDim.x, K + RED_EXTRA, arrays.red, s_tile);
      grid_rendezvous(ctl, epoch, [&] { em_ctl_step<SP_NT>(arrays, ctl, K, 1, s_blk); });
    } else {
      grid_rendezvous(ctl, epoch, [&] {
        reduce_partials_cta<SP_NT>(partials, pstride, (int)gridDim.x, K + RED_EXTRA, arrays.red, arrays.seg);
        __syncthreads();
        em_ctl_step<SP_NT>(arrays, ctl, K, 1, s_blk);
      });
    }
  }
}
inline size_t em_sparse_smem_bytes(int K) {
  return (size_t)K * 16 + (size_t)(SP_NT / 32) * SP_STAGE * (8 + 4) + (size_t)(SP_NT / 32) * 32 * 8;
}

// ---- separate reduction (+ control) kernel: large grids x many groups ---------------------------------
// red[v] = sum over CTAs of partials[cta][v], fixed order; v < nvals.  ctl_mode: -1 none (several GPUs: the all-reduce
// and a control kernel follow), else the last CTA takes the control step (0 EM dense, 1 EM sparse, 2 RCG stage 0).
constexpr int FIN_NT = 1024;   // the control step behind the reduction is K-sized lgamma / digamma / exp work in ONE CTA: 1024 threads take it from ~13 to ~4 us at K = 2000
// peer != 0 (several GPUs with peer memory, peer.cuh): the last CTA also exchanges the reduced vector with the other ranks
// and takes the control step — pass + this kernel are the whole iteration, as on one GPU.
static __global__ void __launch_bounds__(FIN_NT)
finalize_ctl_kernel(const double *partials, int pstride, int n_ctas, int nvals, ViArrays arrays, ViCtl *ctl, int K,
                    int ctl_mode, int ignore_stall, int peer, PeerView pv) {
  if (ctl->done || (ctl->stall && !ignore_stall)) return;
  __shared__ double s_blk[32];
  __shared__ double s_tile[FIN_NT];
  reduce_partials_tiled<FIN_NT>(partials, pstride, n_ctas, nvals, arrays.red, s_tile);
  if (ctl_mode < 0) return;
  if (!cta_is_last(ctl)) return;
  if (threadIdx.x == 0) ctl->ticket = 0;
  __syncthreads();
  if (peer) {
    if (!peer_allreduce_cta<FIN_NT>(arrays.red, nvals, pv)) {
      if (threadIdx.x == 0) { ctl->fault = 2; ctl->done = 1; }
      return;
    }
  }
  if (ctl_mode <= 1) em_ctl_step<FIN_NT>(arrays, ctl, K, ctl_mode, s_blk);
  else rcg_ctl_b_step<FIN_NT>(arrays, ctl, K, 0, peer ? 1 : 0, s_blk);
}

} // namespace mswb
