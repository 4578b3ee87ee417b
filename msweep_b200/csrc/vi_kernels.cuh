// vi_kernels.cuh — the fused VI sweeps (K6 / K7 of SURVEY.md §2.4), EC-major, HBM-bound.
//
// Every sweep has the same shape.  A CTA of NT threads is split into G = NT/TPR row groups; a group of
// TPR threads owns R consecutive classes (rows) per batch and strides the K columns of a row in
// 128-bit pieces: thread t of the group holds columns VEC*(t + TPR*i) .. +VEC-1, i < KITER.  Per-class
// quantities (the logsumexp / the normaliser S_j / the mean step) are reductions ALONG a row:
// warp shuffles, plus one shared-memory hop when a row spans several warps.  Per-group quantities
// (the expected counts N_k) accumulate DOWN the rows in registers and leave the CTA once, as a
// per-CTA partial vector that a second tiny kernel sums in a fixed order (no atomics: results are
// bit-reproducible and identical however many CTAs ran).  The grid is persistent (one CTA per SM, or
// a small multiple) and walks the row batches with a grid stride.
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
};

template <int TPR_, int KITER_, int R_> struct Tile {
  static constexpr int TPR = TPR_, KITER = KITER_, R = R_;
  static constexpr int NT = TPR < 256 ? (256 / TPR) * TPR : TPR;   // as many whole row groups as fit in 256 threads
  static constexpr int G = NT / TPR;
  static constexpr int NW = NT / 32;
  static constexpr int WPG = TPR / 32;   // warps per row group
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

// ---- per-CTA partial vector: fold the G row groups in group order, then one coalesced store ---------
// acc[i][v] belongs to column VEC*(t + TPR*i) + v.  comb: TPR*KITER*VEC doubles when G > 1.
template <class TL, int VEC>
__device__ __forceinline__ void store_partials(const double (&acc)[TL::KITER][VEC], double *comb, double *out, int K) {
  const int t = threadIdx.x % TL::TPR, g = threadIdx.x / TL::TPR;
  if constexpr (TL::G == 1) {
#pragma unroll
    for (int i = 0; i < TL::KITER; ++i)
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const int k = VEC * (t + TL::TPR * i) + v;
        if (k < K) out[k] = acc[i][v];
      }
  } else {
    for (int gg = 0; gg < TL::G; ++gg) {      // fixed order: bit-reproducible
      __syncthreads();
      if (g == gg) {
#pragma unroll
        for (int i = 0; i < TL::KITER; ++i)
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            const int c = VEC * (t + TL::TPR * i) + v;
            comb[c] = gg == 0 ? acc[i][v] : comb[c] + acc[i][v];
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
// EM / VB pass, linear domain.  P(j,k) = exp(logl(j,k) - M_j) is stored once; a pass is two GEMVs that
// share one read of P:   S_j = sum_k P(j,k) w_k ,  A_k = sum_j P(j,k) c_j / S_j ,  with
// w_k = exp(digamma(N_k) - max digamma).  Then N_k = alpha0_k + w_k A_k and the data term of the ELBO
// is sum_j c_j (log S_j + M_j) (+ constants applied by the control kernel).  Two FMAs per element.
//
// P, counts and rowmax are padded with zero rows to a multiple of 64 classes (ROW_PAD): a padded class
// has c = 0 and drops out, so the sweep carries no row-bounds predicates at all.
// =====================================================================================================
constexpr int ROW_PAD = 64;

// Sum R per-row partials over the 32 lanes with a transposing butterfly: 6 (R = 4), 6 (R = 2) or
// 5 (R = 1) 64-bit shuffles instead of 5 R.  On return lane L holds in the result the warp total of
// row rid(L); the lanes with (L & 7) == 0 are the canonical holders (row = holder_row(L)).
template <int R> __device__ __forceinline__ double warp_rows_sum(const double (&s)[R], int lane, int &rid) {
  double k;
  if constexpr (R == 4) {
    const bool hi = lane & 16, h8 = lane & 8;
    double a0 = hi ? s[2] : s[0], a1 = hi ? s[3] : s[1];
    const double b0 = hi ? s[0] : s[2], b1 = hi ? s[1] : s[3];
    a0 += __shfl_xor_sync(0xffffffffu, b0, 16);
    a1 += __shfl_xor_sync(0xffffffffu, b1, 16);
    k = h8 ? a1 : a0;
    const double snd = h8 ? a0 : a1;
    k += __shfl_xor_sync(0xffffffffu, snd, 8);
    rid = (hi ? 2 : 0) + (h8 ? 1 : 0);
  } else if constexpr (R == 2) {
    const bool hi = lane & 16;
    k = hi ? s[1] : s[0];
    const double snd = hi ? s[0] : s[1];
    k += __shfl_xor_sync(0xffffffffu, snd, 16);
    k += __shfl_xor_sync(0xffffffffu, k, 8);
    rid = hi ? 1 : 0;
  } else {
    static_assert(R == 1, "R must be 1, 2 or 4");
    k = s[0];
    k += __shfl_xor_sync(0xffffffffu, k, 16);
    k += __shfl_xor_sync(0xffffffffu, k, 8);
    rid = 0;
  }
  k += __shfl_xor_sync(0xffffffffu, k, 4);
  k += __shfl_xor_sync(0xffffffffu, k, 2);
  k += __shfl_xor_sync(0xffffffffu, k, 1);
  return k;
}
// lane that canonically holds row r after warp_rows_sum<R>
template <int R> __device__ __forceinline__ int holder_lane(int r) { return R == 4 ? 16 * (r >> 1) + 8 * (r & 1) : (R == 2 ? 16 * r : 0); }

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

template <typename ST, class TL, bool PIPE>
__global__ void __launch_bounds__(TL::NT, em_min_blocks<ST, TL>())
em_lin_pass_kernel(const ST *__restrict__ P, int ld, const double *__restrict__ rowmax, const double *__restrict__ counts,
                   const double *__restrict__ w, const ViCtl *__restrict__ ctl, double *__restrict__ partials,
                   int pstride, unsigned long long N_pad, int K, PipeGeom geom) {
  using VT = typename VecOf<ST>::type;
  constexpr int VEC = VecOf<ST>::VEC;
  constexpr int R = TL::R, KITER = TL::KITER, TPR = TL::TPR, WPG = TL::WPG;
  constexpr int RB = TL::G * R;                       // rows per CTA batch
  if (ctl->done) return;
  __shared__ double s_red[2 * TL::NW * R];
  __shared__ double s_comb[TL::G > 1 ? TPR * KITER * VEC : 1];
  __shared__ double s_blk[32];
  const int t = threadIdx.x % TPR, g = threadIdx.x / TPR;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
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

  // ---- one batch: pv = this thread's pieces of R rows, c_lane = count of row (row0 + lane) for lane < R
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
    int rid;
    double k = warp_rows_sum<R>(s, lane, rid);
    double tot;
    if constexpr (WPG > 1) {
      double *buf = s_red + phase * (TL::NW * R);
      phase ^= 1;
      if ((lane & (R == 4 ? 7 : (R == 2 ? 15 : 31))) == 0) buf[warp * R + rid] = k;
      group_sync<TL>(g);
      tot = 0.0;
      if (lane < R) {
        const int w0 = (warp / WPG) * WPG;
#pragma unroll
        for (int ww = 0; ww < WPG; ++ww) tot += buf[(w0 + ww) * R + lane];
      }
    } else {
      // lane r fetches the total of row r from its holder
      const int src = R == 4 ? 16 * ((lane >> 1) & 1) + 8 * (lane & 1) : (R == 2 ? 16 * (lane & 1) : 0);
      tot = __shfl_sync(0xffffffffu, k, src);
    }
    // lanes 0..R-1: row r = lane.  One division per warp serves the whole batch.
    double inv_l = 0.0;
    if (lane < R && c_lane > 0.0) {
      if (!(tot > 0.0) || isinf(tot)) fault = 1;
      else {
        inv_l = c_lane / tot;
        if (t < 32) elbo += c_lane * (log(tot) + rowmax[row0 + lane]);
      }
    }
    ST inv[R];
#pragma unroll
    for (int r = 0; r < R; ++r) inv[r] = (ST)__shfl_sync(0xffffffffu, inv_l, r);
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
      c_lane = lane < R ? counts[row0 + lane] : 0.0;
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
        process(pv, lane < R ? counts[row0 + lane] : 0.0, row0);
      }
    });
  }
  double *out = partials + (unsigned long long)blockIdx.x * pstride;
  store_partials<TL, VEC>(acc, s_comb, out, K);
  elbo = block_sum<TL::NT>(elbo, s_blk);
  if (threadIdx.x == 0) { out[K] = elbo; out[K + 1] = 0.0; }
  if (fault) atomicExch(const_cast<int *>(&ctl->fault), 1);
}

// =====================================================================================================
// Log-domain sweeps (fp64): the RCG optimiser, the restart step, and the log-domain EM pass.
// =====================================================================================================

// Reduce R per-row values over the TPR threads of a row group (sum or max).  On return lanes 0..R-1 of every
// warp of the group hold the group-wide result of row `lane`; other lanes hold unspecified values.
// Transposing butterfly inside the warp (6 / 6 / 5 64-bit shuffles for R = 4 / 2 / 1), then one
// shared-memory hop when the row spans several warps.  scratch: 2 * NW * R doubles.
template <bool IS_MAX> __device__ __forceinline__ double red_op(double a, double b) { return IS_MAX ? fmax(a, b) : a + b; }

template <class TL, bool IS_MAX>
__device__ __forceinline__ double rows_reduce(const double (&v)[TL::R], double *scratch, int &phase, int lane, int warp) {
  constexpr int R = TL::R;
  static_assert(R == 1 || R == 2 || R == 4, "rows per batch must be 1, 2 or 4");
  double k;
  int rid;
  if constexpr (R == 4) {
    const bool hi = lane & 16, h8 = lane & 8;
    double a0 = hi ? v[2] : v[0], a1 = hi ? v[3] : v[1];
    const double b0 = hi ? v[0] : v[2], b1 = hi ? v[1] : v[3];
    a0 = red_op<IS_MAX>(a0, __shfl_xor_sync(0xffffffffu, b0, 16));
    a1 = red_op<IS_MAX>(a1, __shfl_xor_sync(0xffffffffu, b1, 16));
    k = h8 ? a1 : a0;
    const double snd = h8 ? a0 : a1;
    k = red_op<IS_MAX>(k, __shfl_xor_sync(0xffffffffu, snd, 8));
    rid = (hi ? 2 : 0) + (h8 ? 1 : 0);
  } else if constexpr (R == 2) {
    const bool hi = lane & 16;
    k = hi ? v[1] : v[0];
    const double snd = hi ? v[0] : v[1];
    k = red_op<IS_MAX>(k, __shfl_xor_sync(0xffffffffu, snd, 16));
    k = red_op<IS_MAX>(k, __shfl_xor_sync(0xffffffffu, k, 8));
    rid = hi ? 1 : 0;
  } else {
    k = v[0];
    k = red_op<IS_MAX>(k, __shfl_xor_sync(0xffffffffu, k, 16));
    k = red_op<IS_MAX>(k, __shfl_xor_sync(0xffffffffu, k, 8));
    rid = 0;
  }
  k = red_op<IS_MAX>(k, __shfl_xor_sync(0xffffffffu, k, 4));
  k = red_op<IS_MAX>(k, __shfl_xor_sync(0xffffffffu, k, 2));
  k = red_op<IS_MAX>(k, __shfl_xor_sync(0xffffffffu, k, 1));
  if constexpr (TL::WPG > 1) {
    double *buf = scratch + phase * (TL::NW * R);
    phase ^= 1;
    if ((lane & (R == 4 ? 7 : (R == 2 ? 15 : 31))) == 0) buf[warp * R + rid] = k;
    group_sync<TL>(warp / TL::WPG);
    double tot = IS_MAX ? -INFINITY : 0.0;
    if (lane < R) {
      const int w0 = (warp / TL::WPG) * TL::WPG;
#pragma unroll
      for (int ww = 0; ww < TL::WPG; ++ww) tot = red_op<IS_MAX>(tot, buf[(w0 + ww) * R + lane]);
    }
    return tot;
  } else {
    const int src = R == 4 ? 16 * ((lane >> 1) & 1) + 8 * (lane & 1) : (R == 2 ? 16 * (lane & 1) : 0);
    return __shfl_sync(0xffffffffu, k, src);
  }
}

// Sweep A of an RCG iteration ("mixt_negnatgrad"): d = logl + (digamma(N_k) - 1) - gamma,
// newnorm = sum_jk q (d - <d>_j) d  with q = exp(gamma), <d>_j = sum_k q d.  Nothing is written: d is
// recomputed by sweep B, which saves 16 B/element of traffic over storing it.
template <class TL, bool PIPE>
__global__ void __launch_bounds__(TL::NT, TL::NT > 256 ? 1 : (TL::R * TL::KITER <= 4 ? 3 : (TL::R * TL::KITER <= 8 ? 2 : 1)))
rcg_sweep_a_kernel(const double *__restrict__ logl, const double *__restrict__ gamma, int ld,
                   const double *__restrict__ dgm1, const ViCtl *__restrict__ ctl, double *__restrict__ partials,
                   int pstride, unsigned long long N, int K, PipeGeom geom) {
  constexpr int R = TL::R, KITER = TL::KITER, TPR = TL::TPR;
  if (ctl->done) return;
  __shared__ double s_red[2 * TL::NW * R];
  __shared__ double s_blk[32];
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
    if (warp_cols_ok && row0 + R <= N) {
      // fast path (warp-uniform): every row and column this warp touches is real — no predicates
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
      const double sr = __shfl_sync(0xffffffffu, tot, r);
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
  if (threadIdx.x == 0) partials[(unsigned long long)blockIdx.x * pstride + K + 1] = nn;
}

// Sweep B of an RCG iteration: step = d (+ beta * oldstep), gamma += step, renormalise every class,
// store gamma and step, accumulate N_k - alpha0 = sum_j c_j q and the data term of the ELBO.
// MODE 0: RCG step.  MODE 1: plain step from the current digamma vector, gamma = normalise(logl + dg)
// (the RCG restart, and the log-domain EM pass); WRITE says whether gamma is stored.
template <class TL, int MODE, bool WRITE, bool PIPE>
__global__ void __launch_bounds__(TL::NT, TL::R * TL::KITER <= 4 && TL::NT <= 256 ? 2 : 1)
rcg_sweep_b_kernel(const double *__restrict__ logl, double *__restrict__ gamma, double *__restrict__ step, int ld,
                   const double *__restrict__ dgv, const double *__restrict__ counts, const ViCtl *__restrict__ ctl,
                   double *__restrict__ partials, int pstride, unsigned long long N, int K, int only_if_reset,
                   PipeGeom geom) {
  constexpr int R = TL::R, KITER = TL::KITER, TPR = TL::TPR;
  if (ctl->done) return;
  if (only_if_reset && !ctl->didreset) return;
  __shared__ double s_red[2 * TL::NW * R];
  __shared__ double s_comb[TL::G > 1 ? TPR * KITER * 2 : 1];
  __shared__ double s_blk[32];
  const int t = threadIdx.x % TPR, g = threadIdx.x / TPR;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool use_old = MODE == 0 && ctl->use_old != 0;
  const double beta_eff = use_old ? ctl->beta : 0.0;   // step = d + beta * oldstep; no direction memory -> beta 0
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

  // One batch = three element-wise phases separated by two row reductions.  The phases exist in two
  // compile-time flavours: FULL (every row and column this WARP touches is real: no predicates, the loads of a
  // batch issue back to back) and the predicated one for ragged edges.  The choice is warp-uniform and the
  // reductions (shuffles with a full mask, __syncthreads) sit outside the flavoured code, in common control flow.
  const bool warp_cols_ok = __all_sync(0xffffffffu, cols_ok) != 0;
  auto body = [&](auto &&load, unsigned long long row0) {   // load(src, r, idx): 0 = logl, 1 = gamma, 2 = old step
    const bool full = warp_cols_ok && row0 + R <= N;        // uniform over the warp (a warp never spans two row groups)
    double l[R][KITER][2], gn[R][KITER][2], m[R];
    double e[R][KITER][2], sum[R], lsum[R], cinv[R];

    // FULL is a literal at both call sites; the lambda is inlined and specialised for it
    auto phase1 = [&](const bool FULL) {
      double st[R][KITER][2];
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int i = 0; i < KITER; ++i) {
          const int idx = t + TPR * i;
          const bool in = FULL || (row0 + r < N && idx < nvec);
          if (in) unpack(load(0, r, idx), l[r][i]); else l[r][i][0] = l[r][i][1] = 0.0;
          if (MODE == 0) { if (in) unpack(load(1, r, idx), gn[r][i]); else gn[r][i][0] = gn[r][i][1] = 0.0; }
        }
      if (MODE == 0 && use_old) {
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int i = 0; i < KITER; ++i) {
            const int idx = t + TPR * i;
            const bool in = FULL || (row0 + r < N && idx < nvec);
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
            const bool on = FULL || (row < N && ok[i][v]);
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
          if (MODE == 0 && (FULL || (row < N && idx < nvec))) st_stream(sp + idx, make_double2(st[r][i][0], st[r][i][1]));
        }
      }
    };
    auto phase3 = [&](const bool FULL) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const unsigned long long row = row0 + r;
        if (!FULL && row >= N) continue;
        double2 *gp = reinterpret_cast<double2 *>(gamma) + row * (unsigned long long)nvec;
#pragma unroll
        for (int i = 0; i < KITER; ++i) {
          const int idx = t + TPR * i;
          if (!FULL && idx >= nvec) continue;
          double gout[2];
#pragma unroll
          for (int v = 0; v < 2; ++v) {
            const double g2 = gn[r][i][v] - lsum[r];     // normalised log-responsibility
            gout[v] = (FULL || ok[i][v]) ? g2 : 0.0;
            const double cq = cinv[r] * e[r][i][v];      // c_j q(j,k); 0 for unobserved classes and padding
            acc[i][v] += cq;
            if (cq > 0.0) bound = fma(cq, l[r][i][v] - g2, bound);
          }
          if (WRITE) st_stream(gp + idx, make_double2(gout[0], gout[1]));
        }
      }
    };

    if (full) phase1(true); else phase1(false);
    const double mt = rows_reduce<TL, true>(m, s_red, phase, lane, warp);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      double mm = __shfl_sync(0xffffffffu, mt, r);
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
    // lanes 0..R-1: one log, one division and one count load per warp serve the whole batch
    double lsum_l = 0.0, cinv_l = 0.0;
    if (lane < R && row0 + lane < N) {
      lsum_l = log(tot);
      cinv_l = counts[row0 + lane] / tot;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      lsum[r] = __shfl_sync(0xffffffffu, lsum_l, r);
      cinv[r] = __shfl_sync(0xffffffffu, cinv_l, r);
    }
    if (full) phase3(true); else phase3(false);
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
  if (threadIdx.x == 0) out[K] = bound;
}

// =====================================================================================================
// EM / VB pass over the SPARSE storage: class j is P0_j on every group except its hits, where it is P0_j + dP.
//   S_j = P0_j W + sum_hits dP w_g          (W = sum_k w_k)
//   A_k = Z + sum_{hits of k} dP c_j / S_j  (Z = sum_j P0_j c_j / S_j, the part every group receives)
// One thread per class, the weight vector and the per-group accumulators in shared memory (fp64 atomics: the
// only place on the path where the summation order, hence the last bits, can vary from run to run).
// =====================================================================================================
__global__ void __launch_bounds__(256)
em_sparse_pass_kernel(const uint64_t *__restrict__ nz_ptr, const uint32_t *__restrict__ nz_grp, const double *__restrict__ nz_dP,
                      const double *__restrict__ P0, const double *__restrict__ rowmax, const double *__restrict__ counts,
                      const double *__restrict__ w, const ViCtl *__restrict__ ctl, double *__restrict__ partials, int pstride,
                      unsigned long long N, int K) {
  if (ctl->done) return;
  extern __shared__ double s_dyn[];
  double *s_w = s_dyn, *s_acc = s_dyn + K;
  __shared__ double s_blk[32];
  double wsum = 0.0;
  for (int k = threadIdx.x; k < K; k += 256) { const double x = w[k]; s_w[k] = x; s_acc[k] = 0.0; wsum += x; }
  wsum = block_sum<256>(wsum, s_blk);        // same order in every CTA and on every rank
  __syncthreads();
  double z = 0.0, elbo = 0.0;
  int fault = 0;
  // The pass is bound by load latency, not by bytes: everything a class needs is requested before anything is used
  // (its five per-class values together, then its first HEAD hits together), and the hits stay in registers for the
  // scatter.  Terms are added in hit order whatever the chunking, so the sums do not depend on HEAD.
  constexpr int HEAD = 8;
  for (unsigned long long j = blockIdx.x * 256ull + threadIdx.x; j < N; j += (unsigned long long)gridDim.x * 256ull) {
    const double c = counts[j];
    const unsigned long long a = nz_ptr[j], b = nz_ptr[j + 1];
    const double p0 = P0[j], m = rowmax[j];
    if (!(c > 0.0)) continue;
    const unsigned long long n = b - a;
    double dp[HEAD];
    uint32_t gr[HEAD];
#pragma unroll
    for (int u = 0; u < HEAD; ++u) {
      const bool on = (unsigned long long)u < n;
      dp[u] = on ? nz_dP[a + u] : 0.0;
      gr[u] = on ? nz_grp[a + u] : 0u;
    }
    double s = p0 * wsum;
#pragma unroll
    for (int u = 0; u < HEAD; ++u) s = fma(dp[u], s_w[gr[u]], s);          // a padded slot adds dp = 0: s unchanged
    for (unsigned long long e = a + HEAD; e < b; ++e) s = fma(nz_dP[e], s_w[nz_grp[e]], s);
    if (!(s > 0.0) || isinf(s)) { fault = 1; continue; }
    const double r = c / s;
    z = fma(r, p0, z);
    elbo = fma(c, log(s) + m, elbo);
#pragma unroll
    for (int u = 0; u < HEAD; ++u)
      if ((unsigned long long)u < n) atomicAdd(&s_acc[gr[u]], r * dp[u]);
    for (unsigned long long e = a + HEAD; e < b; ++e) atomicAdd(&s_acc[nz_grp[e]], r * nz_dP[e]);
  }
  __syncthreads();
  double *out = partials + (unsigned long long)blockIdx.x * pstride;
  for (int k = threadIdx.x; k < K; k += 256) out[k] = s_acc[k];
  z = block_sum<256>(z, s_blk);
  elbo = block_sum<256>(elbo, s_blk);
  if (threadIdx.x == 0) { out[K] = elbo; out[K + 1] = z; }
  if (fault) atomicExch(const_cast<int *>(&ctl->fault), 1);
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
