// bootstrap.cu — BootstrapSample on the device (src/BootstrapSample.cpp:33-73 and the replicate loop
// src/mSWEEP.cpp:496-518 of the reference).
//
// The reference draws `bootstrap_count` categorical samples one at a time from
// std::discrete_distribution<uint32_t>(class counts) driven by std::mt19937_64, i.e. per draw:
// one 64-bit Mersenne-Twister output u, p = double(u) * 2^-64 (nextafter(1,0) if that rounds to 1),
// class = lower_bound(cumulative probabilities, p).  In MSWB_RNG_LIBSTDCXX_EXACT mode the very same generator
// stream is produced ON THE DEVICE: the MT19937-64 recurrence x[k+312] = x[k+156] ^ f(x[k], x[k+1]) reaches back at
// least 156 words, so one CTA advances the state 156 words at a time in parallel (mt64_generate_kernel: ~3 ms for the 1e7
// draws of a replicate, against ~80 ms for std::mt19937_64 on a host core plus the copy); the binary searches over the N
// cumulative probabilities and the histogram follow, and the resampled counts are bit-identical to the reference's.
// MSWB_RNG_PHILOX replaces the stream by a counter-based generator.
#include "handles.cuh"

#include <cmath>
#include <cstdlib>
#include <memory>
#include <random>

using namespace mswb;

namespace mswb {

__device__ __forceinline__ double canonical_from_u64(unsigned long long u) {
  // std::generate_canonical<double, 53>(mt19937_64): one output, rounded to nearest, times 2^-64
  double p = __ull2double_rn(u) * 0x1p-64;
  return p >= 1.0 ? 0x1.fffffffffffffp-1 : p;
}

__device__ __forceinline__ unsigned long long lower_bound_idx(const double *__restrict__ cp, unsigned long long n, double p) {
  unsigned long long lo = 0, hi = n;   // first index with cp[idx] >= p
  while (lo < hi) {
    const unsigned long long mid = (lo + hi) >> 1;
    if (cp[mid] < p) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void resample_from_stream_kernel(const unsigned long long *__restrict__ stream, unsigned long long n_draws,
                                            const double *__restrict__ cp, unsigned long long n_cp,
                                            unsigned *__restrict__ hist) {
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n_draws;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long idx = n_cp ? lower_bound_idx(cp, n_cp, canonical_from_u64(stream[i])) : 0ull;
    atomicAdd(&hist[idx], 1u);
  }
}

// ---- std::mt19937_64 on the device ----------------------------------------------------------------------------------
// state[0..311] = the generator's words, state[312] = how many of them have been consumed (312: refill first, the state
// std::mt19937_64 is in right after seeding).  Produces the next n outputs of the stream into out (nullptr: discard — the
// replicates of other ranks) and leaves the state where the stream continues.
constexpr int MT_N = 312, MT_M = 156, MT_NT = 160;
__device__ __forceinline__ unsigned long long mt64_temper(unsigned long long y) {
  y ^= (y >> 29) & 0x5555555555555555ull;
  y ^= (y << 17) & 0x71D67FFFEDA60000ull;
  y ^= (y << 37) & 0xFFF7EEE000000000ull;
  y ^= (y >> 43);
  return y;
}
__device__ __forceinline__ unsigned long long mt64_twist(unsigned long long far, unsigned long long cur, unsigned long long nxt) {
  const unsigned long long y = (cur & 0xFFFFFFFF80000000ull) | (nxt & 0x7FFFFFFFull);
  return far ^ (y >> 1) ^ ((y & 1ull) ? 0xB5026F5AA96619E9ull : 0ull);
}
__global__ void __launch_bounds__(MT_NT) mt64_generate_kernel(unsigned long long *__restrict__ state, unsigned long long n,
                                                              unsigned long long *__restrict__ out) {
  // two copies of the state: a refill reads one and writes the other, so each half-step needs ONE barrier
  __shared__ unsigned long long xs[2][MT_N];
  const int t = threadIdx.x;
  for (int i = t; i < MT_N; i += MT_NT) xs[0][i] = state[i];
  unsigned long long pos = state[MT_N];
  int cur = 0;
  __syncthreads();
  unsigned long long produced = 0;
  while (produced < n) {
    if (pos == MT_N) {
      const unsigned long long *x = xs[cur];
      unsigned long long *y = xs[cur ^ 1];
      // words 0..155 depend on old words only; words 156..311 on the NEW first half and old words (word 311 on new word 0)
      if (t < MT_M) y[t] = mt64_twist(x[t + MT_M], x[t], x[t + 1]);
      __syncthreads();
      if (t < MT_M) y[t + MT_M] = mt64_twist(y[t], x[t + MT_M], t + MT_M + 1 < MT_N ? x[t + MT_M + 1] : y[0]);
      __syncthreads();
      cur ^= 1;
      pos = 0;
    }
    const unsigned long long take = min((unsigned long long)MT_N - pos, n - produced);
    if (out) for (unsigned long long i = t; i < take; i += MT_NT) out[produced + i] = mt64_temper(xs[cur][pos + i]);
    produced += take;
    pos += take;
  }
  __syncthreads();
  for (int i = t; i < MT_N; i += MT_NT) state[i] = xs[cur][i];
  if (t == 0) state[MT_N] = pos;
}

// ---- the same stream in PARALLEL segments ---------------------------------------------------------------------------
// One CTA advances the generator at ~6e8 outputs a second: the 1e7 draws of a config-5 replicate take 18 ms, about as long
// as the optimiser run that consumes them.  MT19937-64 is linear over GF(2) (mt64_jump.cu): the state L outputs ahead is
// g_L(A) applied to the state, g_L = z^L mod phi computed once per bootstrap run on the host.  So a replicate's draws are
// cut into segments of L outputs (L a multiple of 312: every segment starts at a refill boundary of the sequential
// generator), mt64_chain_kernel walks the segment starts B_0, B_1 = g_L(A) B_0, ... (one CTA: 19937 steps of the recurrence
// into shared memory, then word m of the next start = XOR of raw[i + m] over the set coefficients i — ~70 us a start) and
// mt64_segments_kernel generates all segments at once, one CTA each.  The outputs and the state left behind are those of
// the sequential generator bit for bit (tests/test_gpu_parity.py::test_bootstrap_counts_bit_exact,
// test_mt64_segments_equal_the_sequential_stream).
constexpr int MT_RAW = 19937 + MT_N;          // raw words the application of a jump polynomial reads
constexpr int MT_CHAIN_NT = 960;              // three groups of 320 threads (312 active) share the set coefficients
constexpr size_t MT_CHAIN_SMEM = (size_t)(MT_RAW + 313 + 3 * 320) * sizeof(unsigned long long);

// state[0..311] + state[312] = consumed count (the sequential generator's state).  First the n_head outputs still waiting in
// the current array are tempered out (the array is then a refill boundary), then starts[s] = the array jumped s segments
// ahead, s = 0 .. n_starts - 1, each with its consumed count set to 312.
__global__ void __launch_bounds__(MT_CHAIN_NT) mt64_chain_kernel(unsigned long long *__restrict__ state, unsigned long long n_head,
                                                                  unsigned long long *__restrict__ out,
                                                                  const unsigned long long *__restrict__ gbits, int n_starts,
                                                                  unsigned long long *__restrict__ starts) {
  extern __shared__ unsigned long long mt_sm[];
  unsigned long long *raw = mt_sm;               // [MT_RAW]
  unsigned long long *g = raw + MT_RAW;          // [313]
  unsigned long long *part = g + 313;            // [3][320]
  const int t = threadIdx.x;
  const unsigned long long pos = state[MT_N];
  for (int i = t; i < 313; i += MT_CHAIN_NT) g[i] = gbits[i];
  for (int i = t; i < MT_N; i += MT_CHAIN_NT) raw[i] = state[i];
  __syncthreads();
  if (out) for (unsigned long long i = t; i < n_head; i += MT_CHAIN_NT) out[i] = mt64_temper(raw[pos + i]);
  if (t == 0) state[MT_N] = pos + n_head;
  for (int s = 0; s < n_starts; ++s) {
    for (int i = t; i < MT_N; i += MT_CHAIN_NT) starts[(size_t)s * (MT_N + 1) + i] = raw[i];
    if (t == 0) starts[(size_t)s * (MT_N + 1) + MT_N] = MT_N;
    if (s + 1 == n_starts) break;
    // the recurrence reaches back at least 156 words: blocks of 156 words in parallel, five warps behind a named barrier
    if (t < 160) {
      for (int k0 = MT_N; k0 < MT_RAW; k0 += MT_M) {
        const int k = k0 + t;
        if (t < MT_M && k < MT_RAW) raw[k] = mt64_twist(raw[k - MT_M], raw[k - MT_N], raw[k - MT_N + 1]);
        asm volatile("bar.sync 1, 160;" ::: "memory");
      }
    }
    __syncthreads();
    const int grp = t / 320, m = t - grp * 320;
    unsigned long long acc = 0ull;
    if (m < MT_N) {
      for (int w = grp; w < 313; w += 3) {
        unsigned long long bits = g[w];
        while (bits) {
          const int i = 64 * w + __ffsll((long long)bits) - 1;
          bits &= bits - 1ull;
          acc ^= raw[i + m];
        }
      }
    }
    part[t] = acc;
    __syncthreads();
    if (t < MT_N) raw[t] = part[t] ^ part[320 + t] ^ part[640 + t];
    __syncthreads();
  }
}

// CTA s: the outputs [s * seg_len, min((s + 1) * seg_len, n)) of the stream that continues from starts[s]; the last CTA leaves
// the generator's state in `state_out`.
__global__ void __launch_bounds__(MT_NT) mt64_segments_kernel(const unsigned long long *__restrict__ starts, unsigned long long seg_len,
                                                             unsigned long long n, unsigned long long *__restrict__ out,
                                                             unsigned long long *__restrict__ state_out) {
  __shared__ unsigned long long xs[2][MT_N];
  const int t = threadIdx.x;
  const unsigned long long first = (unsigned long long)blockIdx.x * seg_len;
  if (first >= n) return;
  const unsigned long long mine = min(seg_len, n - first);
  const unsigned long long *st = starts + (size_t)blockIdx.x * (MT_N + 1);
  for (int i = t; i < MT_N; i += MT_NT) xs[0][i] = st[i];
  unsigned long long pos = MT_N;
  int cur = 0;
  __syncthreads();
  unsigned long long produced = 0;
  out += first;
  while (produced < mine) {
    if (pos == MT_N) {
      const unsigned long long *x = xs[cur];
      unsigned long long *y = xs[cur ^ 1];
      if (t < MT_M) y[t] = mt64_twist(x[t + MT_M], x[t], x[t + 1]);
      __syncthreads();
      if (t < MT_M) y[t + MT_M] = mt64_twist(y[t], x[t + MT_M], t + MT_M + 1 < MT_N ? x[t + MT_M + 1] : y[0]);
      __syncthreads();
      cur ^= 1;
      pos = 0;
    }
    const unsigned long long take = min((unsigned long long)MT_N - pos, mine - produced);
    for (unsigned long long i = t; i < take; i += MT_NT) out[produced + i] = mt64_temper(xs[cur][pos + i]);
    produced += take;
    pos += take;
  }
  if (first + mine == n) {
    __syncthreads();
    for (int i = t; i < MT_N; i += MT_NT) state_out[i] = xs[cur][i];
    if (t == 0) state_out[MT_N] = pos;
  }
}

// Philox-4x32-10 keyed by (seed, replicate), counter = draw index / 2; two 64-bit outputs per block.
__device__ __forceinline__ void philox_round(unsigned (&c)[4], unsigned (&k)[2]) {
  const unsigned long long p0 = 0xD2511F53ull * c[0], p1 = 0xCD9E8D57ull * c[2];
  const unsigned n0 = (unsigned)(p1 >> 32) ^ c[1] ^ k[0], n1 = (unsigned)p1;
  const unsigned n2 = (unsigned)(p0 >> 32) ^ c[3] ^ k[1], n3 = (unsigned)p0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
}
__global__ void resample_philox_kernel(unsigned long long seed, unsigned long long replicate, unsigned long long n_draws,
                                       const double *__restrict__ cp, unsigned long long n_cp, unsigned *__restrict__ hist) {
  const unsigned long long n_blocks = (n_draws + 1) / 2;
  for (unsigned long long b = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; b < n_blocks;
       b += (unsigned long long)gridDim.x * blockDim.x) {
    unsigned c[4] = {(unsigned)b, (unsigned)(b >> 32), (unsigned)replicate, (unsigned)(replicate >> 32)};
    unsigned k[2] = {(unsigned)seed, (unsigned)(seed >> 32)};
#pragma unroll
    for (int r = 0; r < 10; ++r) philox_round(c, k);
    const unsigned long long u0 = ((unsigned long long)c[1] << 32) | c[0], u1 = ((unsigned long long)c[3] << 32) | c[2];
    atomicAdd(&hist[n_cp ? lower_bound_idx(cp, n_cp, canonical_from_u64(u0)) : 0ull], 1u);
    if (2 * b + 1 < n_draws) atomicAdd(&hist[n_cp ? lower_bound_idx(cp, n_cp, canonical_from_u64(u1)) : 0ull], 1u);
  }
}

__global__ void u32_to_double_kernel(const unsigned *in, double *out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = (double)in[i];
}

} // namespace mswb

// declared in vi.cu: a run whose class counts already sit on the device
int mswb_vi_run_dev_counts(mswb_ctx *ctx, mswb_lik *lik, const double *alpha0, const double *counts_dev,
                           double sum_counts, const mswb_vi_opts *opts, double *theta, mswb_vi_stat *stat);
// ... and the batched form: B count vectors, one sweep of the likelihood per iteration for all of them (vi.cu)
bool mswb_vi_batch_supported(const mswb_lik *lik, const mswb_vi_opts *opts);
int mswb_vi_run_batch_dev_counts(mswb_ctx *ctx, mswb_lik *lik, const double *alpha0, const double *counts_dev,
                                 uint64_t counts_stride, int B, double sum_counts, const mswb_vi_opts *opts, double *thetas,
                                 mswb_vi_stat *stats);

namespace mswb {      // mt64_jump.cu
bool mt64_jump_available();
void mt64_jump(const uint64_t *state, uint64_t n_outputs, uint64_t *out);
void mt64_jump_poly(uint64_t n_outputs, uint64_t *bits_out);
}

namespace {

struct Resampler {
  mswb_ctx *ctx;
  uint64_t N = 0, draws = 0;
  int rng_mode;
  uint64_t seed64 = 0;
  // Jump mode (replicates spread over ranks): the host keeps the generator's words at the start of replicate host_rep and
  // moves them with mt64_jump — the draws of the other ranks' replicates are never produced.
  bool jump_mode = false;
  std::vector<unsigned long long> host_win;
  uint64_t host_rep = 0;
  DevBuf<double> cp;           // cumulative probabilities (empty when N < 2, as in libstdc++)
  uint64_t n_cp = 0;
  DevBuf<unsigned long long> mt_state;     // [313] std::mt19937_64's words + consumed count (mt64_generate_kernel)
  DevBuf<unsigned long long> stream_dev;
  static constexpr uint64_t CHUNK = 1u << 24;   // draws generated and consumed at a time (128 MB of stream)
  // parallel segments (mt64_chain_kernel / mt64_segments_kernel): used when a replicate has enough draws to pay for the
  // jump polynomial (~50 ms of host time, once per run) and the chain; MSWB_MT_SEGMENTS=0 / n overrides
  int n_segments = 0;
  uint64_t seg_len = 0;
  uint64_t host_pos = MT_N;                     // consumed count of the device generator, tracked on the host
  DevBuf<unsigned long long> seg_poly, seg_starts;

  Resampler(mswb_ctx *c, const mswb_lik *lik, int32_t seed, uint64_t bootstrap_count, int mode) : ctx(c), rng_mode(mode) {
    MSWB_REQUIRE(mode == MSWB_RNG_LIBSTDCXX_EXACT || mode == MSWB_RNG_PHILOX, "unknown rng mode");
    MSWB_REQUIRE(lik->N == lik->N_total, "bootstrap needs the whole class table on this GPU (use a world_size == 1 context per replica)");
    N = lik->N;
    cudaStream_t s = ctx->stream;
    // src/BootstrapSample.cpp:38-44: uint32_t weights -> libstdc++ discrete_distribution::_M_initialize
    std::vector<double> c_host(N);
    d2h(c_host.data(), lik->counts.p, N, s);
    MSWB_CUDA(cudaStreamSynchronize(s));
    uint64_t total = 0;
    std::vector<double> prob(N);
    for (uint64_t i = 0; i < N; ++i) { const uint32_t w = (uint32_t)c_host[i]; prob[i] = (double)w; total += (uint64_t)c_host[i]; }
    draws = bootstrap_count == 0 ? total : bootstrap_count;          // src/BootstrapSample.cpp:56
    if (N >= 2) {
      double sum = 0.0;
      for (uint64_t i = 0; i < N; ++i) sum += prob[i];
      MSWB_REQUIRE(sum > 0.0, "all class counts are zero");
      for (uint64_t i = 0; i < N; ++i) prob[i] /= sum;
      double run = 0.0;
      for (uint64_t i = 0; i < N; ++i) { run += prob[i]; prob[i] = run; }   // std::partial_sum
      prob[N - 1] = 1.0;
      cp.alloc(N);
      h2d(cp.p, prob.data(), N, s);
      MSWB_CUDA(cudaStreamSynchronize(s));
      n_cp = N;
    }
    // src/BootstrapSample.cpp:48-53 (seed narrowed to int32_t by include/Sample.hpp:163-169)
    if (seed == 26012023) { std::random_device rd; seed64 = rd(); }
    else seed64 = (uint64_t)(int64_t)seed;
    if (mode == MSWB_RNG_LIBSTDCXX_EXACT) {
      // std::mersenne_twister_engine<uint64_t, 64, 312, ...>::seed(value): x[0] = value, x[i] = f * (x[i-1] ^ (x[i-1] >> 62)) + i,
      // and the first output refills
      std::vector<unsigned long long> st(MT_N + 1);
      st[0] = seed64;
      for (int i = 1; i < MT_N; ++i) st[i] = 6364136223846793005ull * (st[i - 1] ^ (st[i - 1] >> 62)) + (unsigned long long)i;
      st[MT_N] = MT_N;
      host_win = st;
      mt_state.alloc(MT_N + 1);
      h2d(mt_state.p, st.data(), st.size(), s);
      MSWB_CUDA(cudaStreamSynchronize(s));
      stream_dev.alloc(std::min<uint64_t>(CHUNK, std::max<uint64_t>(draws, 1)));
    }
  }

  // Decide whether the stream of a replicate is generated in parallel segments; n_mine = replicates this rank resamples.
  void choose_segments(uint64_t n_mine) {
    if (rng_mode != MSWB_RNG_LIBSTDCXX_EXACT || draws == 0) return;
    int want = 16;
    if (const char *e = std::getenv("MSWB_MT_SEGMENTS")) want = std::atoi(e);
    else if ((double)draws * (double)n_mine < 4e7) want = 0;          // the polynomial would cost more than it saves
    if (want < 2 || !mt64_jump_available()) return;
    want = std::min(want, 64);
    const uint64_t per_chunk = std::min<uint64_t>(CHUNK, draws);
    seg_len = (uint64_t)MT_N * ceil_div(per_chunk, (uint64_t)MT_N * (uint64_t)want);
    std::vector<unsigned long long> bits(313);
    mt64_jump_poly(seg_len, reinterpret_cast<uint64_t *>(bits.data()));
    seg_poly.alloc(313);
    h2d(seg_poly.p, bits.data(), bits.size(), ctx->stream);
    MSWB_CUDA(cudaStreamSynchronize(ctx->stream));
    seg_starts.alloc((size_t)want * (MT_N + 1));
    MSWB_CUDA(cudaFuncSetAttribute(mt64_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MT_CHAIN_SMEM));
    n_segments = want;
  }

  static uint64_t pos_after(uint64_t pos, uint64_t n) {      // consumed count of the array after n more outputs
    if (n <= (uint64_t)MT_N - pos) return pos + n;
    const uint64_t rest = n - ((uint64_t)MT_N - pos);
    return (rest - 1) % (uint64_t)MT_N + 1;
  }

  // the next n outputs of the stream into stream_dev (n <= CHUNK)
  void generate(uint64_t n) {
    cudaStream_t s = ctx->stream;
    if (n_segments >= 2 && n > (uint64_t)MT_N) {
      const uint64_t head = std::min<uint64_t>(n, (uint64_t)MT_N - host_pos);
      const uint64_t rest = n - head;
      const int n_starts = (int)ceil_div(rest, seg_len);
      MSWB_REQUIRE(n_starts <= n_segments, "segment plan does not cover the chunk");
      mt64_chain_kernel<<<1, MT_CHAIN_NT, MT_CHAIN_SMEM, s>>>(mt_state.p, head, stream_dev.p, seg_poly.p, n_starts, seg_starts.p);
      MSWB_LAUNCHED();
      if (n_starts > 0) {
        mt64_segments_kernel<<<n_starts, MT_NT, 0, s>>>(seg_starts.p, seg_len, rest, stream_dev.p + head, mt_state.p);
        MSWB_LAUNCHED();
      }
    } else {
      mt64_generate_kernel<<<1, MT_NT, 0, s>>>(mt_state.p, n, stream_dev.p);
      MSWB_LAUNCHED();
    }
    host_pos = pos_after(host_pos, n);
  }

  // Replicates of other ranks ahead: jump over them when producing their draws would cost more than the two polynomials
  // a run needs (about 0.1 s of host time; the generator makes ~6e8 draws a second).  MSWB_MT_JUMP=0 / 1 overrides.
  void choose_jump(uint64_t n_replicates, int replica_rank, int replica_world) {
    if (rng_mode != MSWB_RNG_LIBSTDCXX_EXACT || draws == 0 || replica_world <= 1) return;
    uint64_t mine = 0;
    for (uint64_t r = 0; r < n_replicates; ++r) mine += (int)(r % (uint64_t)replica_world) == replica_rank;
    const char *e = std::getenv("MSWB_MT_JUMP");
    const bool want = e ? std::atoi(e) != 0 : (double)draws * (double)(n_replicates - mine) >= 5e7;
    if (want && (double)draws * (double)n_replicates < 1.8e19 && mt64_jump_available()) { jump_mode = true; host_rep = 0; }
  }

  // another rank's replicate: its draws are consumed, nothing else
  void skip() {
    if (rng_mode != MSWB_RNG_LIBSTDCXX_EXACT || draws == 0 || jump_mode) return;
    mt64_generate_kernel<<<1, MT_NT, 0, ctx->stream>>>(mt_state.p, draws, nullptr);
    MSWB_LAUNCHED();
    host_pos = pos_after(host_pos, draws);
  }

  // hist_dev[N] (zeroed here) receives the resampled class counts of the next replicate; everything is enqueued, nothing waits
  void next(uint64_t replicate, unsigned *hist_dev) {
    cudaStream_t s = ctx->stream;
    MSWB_CUDA(cudaMemsetAsync(hist_dev, 0, std::max<uint64_t>(N, 1) * sizeof(unsigned), s));
    const int grid = ctx->n_sms * 8;
    if (rng_mode == MSWB_RNG_PHILOX) {
      if (draws) { resample_philox_kernel<<<grid, 256, 0, s>>>(seed64, replicate, draws, cp.p, n_cp, hist_dev); MSWB_LAUNCHED(); }
      return;
    }
    if (jump_mode) {
      MSWB_REQUIRE(replicate >= host_rep, "bootstrap replicates must be taken in ascending order");
      if (replicate != host_rep) {
        std::vector<unsigned long long> moved(MT_N + 1);
        mt64_jump(reinterpret_cast<const uint64_t *>(host_win.data()), (replicate - host_rep) * draws, reinterpret_cast<uint64_t *>(moved.data()));
        moved[MT_N] = MT_N;
        host_win.swap(moved);
        host_rep = replicate;
      }
      h2d(mt_state.p, host_win.data(), host_win.size(), s);      // (pageable source: staged before the call returns)
      host_pos = MT_N;
    }
    for (uint64_t done = 0; done < draws; done += CHUNK) {
      const uint64_t n = std::min<uint64_t>(CHUNK, draws - done);
      generate(n);
      resample_from_stream_kernel<<<grid, 256, 0, s>>>(stream_dev.p, n, cp.p, n_cp, hist_dev);
      MSWB_LAUNCHED();
    }
  }
};

} // namespace

extern "C" {

int mswb_bootstrap_resample(mswb_ctx *ctx, const mswb_lik *lik, int32_t seed, uint64_t bootstrap_count,
                            int rng_mode, uint64_t n_replicates, uint32_t *out) {
  return guarded([&] {
    MSWB_REQUIRE(ctx && lik && out, "NULL argument");
    MSWB_CUDA(cudaSetDevice(ctx->device));
    Resampler rs(ctx, lik, seed, bootstrap_count, rng_mode);
    rs.choose_segments(n_replicates);
    DevBuf<unsigned> hist;
    hist.alloc(rs.N);
    for (uint64_t r = 0; r < n_replicates; ++r) {
      rs.next(r, hist.p);
      d2h((unsigned *)out + r * rs.N, hist.p, rs.N, ctx->stream);
      MSWB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
  });
}

int mswb_bootstrap_run(mswb_ctx *ctx, mswb_lik *lik, const double *alpha0, const mswb_vi_opts *opts,
                       uint64_t n_replicates, uint64_t bootstrap_count, int32_t seed, int rng_mode,
                       int replica_rank, int replica_world, double *thetas, mswb_vi_stat *stats) {
  return guarded([&] {
    MSWB_REQUIRE(ctx && lik && alpha0 && opts && thetas, "NULL argument");
    MSWB_REQUIRE(replica_world >= 1 && replica_rank >= 0 && replica_rank < replica_world, "bad replica rank / world");
    MSWB_CUDA(cudaSetDevice(ctx->device));
    Resampler rs(ctx, lik, seed, bootstrap_count, rng_mode);
    rs.choose_jump(n_replicates, replica_rank, replica_world);
    {
      uint64_t n_mine = 0;
      for (uint64_t r = 0; r < n_replicates; ++r) n_mine += (int)(r % (uint64_t)replica_world) == replica_rank;
      rs.choose_segments(n_mine);
    }
    DevBuf<unsigned> hist;
    DevBuf<double> counts;
    hist.alloc(rs.N);
    // EM on a dense likelihood: the replicates of this rank run as device batches — all their count vectors are
    // resampled first, then one sweep of the matrix per iteration serves every replicate still running (vi.cu).
    // Groups of at most 64 replicates and 4 GB of counts at a time.
    if (mswb_vi_batch_supported(lik, opts)) {
      std::vector<uint64_t> mine;
      for (uint64_t r = 0; r < n_replicates; ++r) if ((int)(r % (uint64_t)replica_world) == replica_rank) mine.push_back(r);
      const uint64_t group_max = std::max<uint64_t>(1, std::min<uint64_t>(64, ((uint64_t)4 << 30) / (lik->N_pad * sizeof(double))));
      uint64_t next_rep = 0;      // next replicate of the sequence whose draws have not been consumed yet
      for (size_t g0 = 0; g0 < mine.size(); g0 += group_max) {
        const size_t nb = std::min<size_t>(group_max, mine.size() - g0);
        counts.alloc(nb * lik->N_pad);
        MSWB_CUDA(cudaMemsetAsync(counts.p, 0, counts.bytes(), ctx->stream));
        for (size_t b = 0; b < nb; ++b) {
          const uint64_t r = mine[g0 + b];
          for (; next_rep < r; ++next_rep) rs.skip();          // another GPU's replicates
          rs.next(r, hist.p);
          ++next_rep;
          u32_to_double_kernel<<<ctx->n_sms * 2, 256, 0, ctx->stream>>>(hist.p, counts.p + b * lik->N_pad, rs.N);
          MSWB_LAUNCHED();
        }
        std::vector<double> th(nb * lik->K);
        std::vector<mswb_vi_stat> st(nb);
        if (mswb_vi_run_batch_dev_counts(ctx, lik, alpha0, counts.p, lik->N_pad, (int)nb, (double)rs.draws, opts, th.data(), st.data()))
          throw Error(std::string("bootstrap replicates ") + std::to_string(mine[g0]) + ".." + std::to_string(mine[g0 + nb - 1]) + ": " + mswb_last_error());
        for (size_t b = 0; b < nb; ++b) {
          std::copy(th.begin() + b * lik->K, th.begin() + (b + 1) * lik->K, thetas + mine[g0 + b] * lik->K);
          if (stats) stats[mine[g0 + b]] = st[b];
        }
      }
      return;
    }
    counts.alloc(lik->N_pad);
    MSWB_CUDA(cudaMemsetAsync(counts.p, 0, counts.bytes(), ctx->stream));
    for (uint64_t r = 0; r < n_replicates; ++r) {
      if ((int)(r % (uint64_t)replica_world) != replica_rank) { rs.skip(); continue; }   // another GPU's replicate
      rs.next(r, hist.p);
      u32_to_double_kernel<<<ctx->n_sms * 2, 256, 0, ctx->stream>>>(hist.p, counts.p, rs.N);
      MSWB_LAUNCHED();
      // the reference feeds log(count) with -inf for unsampled classes (src/BootstrapSample.cpp:67-72);
      // the sweeps multiply by the count itself, so a zero count simply drops the class.
      if (mswb_vi_run_dev_counts(ctx, lik, alpha0, counts.p, (double)rs.draws, opts, thetas + r * lik->K,
                                 stats ? stats + r : nullptr))
        throw Error(std::string("bootstrap replicate ") + std::to_string(r) + ": " + mswb_last_error());
    }
  });
}

} // extern "C"
