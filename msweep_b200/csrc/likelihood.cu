// likelihood.cu — LL_WOR21 on the device (include/Likelihood.hpp:92-207 of the reference):
// per-(class, group) hit counts from the class patterns, the --min-hits mask, the beta-binomial
// lookup table and the dense gather, written EC-major straight into the layout the VI sweeps read.
#include "handles.cuh"

#include <cub/cub.cuh>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>

using namespace mswb;

namespace mswb {

// One warp owns one class at a time and a private row of K_all counters in shared memory.
// Counting is O(pattern length); the counters are zeroed again by walking the same pattern, so the
// O(K_all) cost per class is paid only by the kernels that have to emit a dense row anyway.
constexpr int LIK_NT = 256;
constexpr int LIK_WARPS = LIK_NT / 32;

__device__ __forceinline__ void count_pattern(unsigned *row, const uint32_t *__restrict__ tg, unsigned long long a,
                                              unsigned long long b, const uint32_t *__restrict__ group_of_target, int lane) {
  for (unsigned long long p = a + lane; p < b; p += 32) atomicAdd(&row[group_of_target[tg[p]]], 1u);
  __syncwarp();
}

// --min-hits tallies: hits[g] += ec_count[i] once per (class, group) with c(g,i) > 0
// (include/Likelihood.hpp:149-154).  atomicExch hands each distinct group to exactly one lane and
// leaves the counter row clean for the next class.
__global__ void __launch_bounds__(LIK_NT)
group_hits_kernel(const uint64_t *__restrict__ pat_ptr, const uint32_t *__restrict__ pat_targets,
                  const uint32_t *__restrict__ group_of_target, const double *__restrict__ counts,
                  unsigned long long N, int K_all, int active_warps, unsigned long long *__restrict__ hits) {
  extern __shared__ unsigned s_rows[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp >= active_warps) return;   // (no block-wide barrier below: warps are independent)
  unsigned *row = s_rows + (size_t)warp * K_all;
  for (int k = lane; k < K_all; k += 32) row[k] = 0;
  __syncwarp();
  const unsigned long long n_warps = (unsigned long long)gridDim.x * active_warps;
  for (unsigned long long j = (unsigned long long)blockIdx.x * active_warps + warp; j < N; j += n_warps) {
    const unsigned long long a = pat_ptr[j], b = pat_ptr[j + 1];
    count_pattern(row, pat_targets, a, b, group_of_target, lane);
    const unsigned long long c = (unsigned long long)counts[j];
    for (unsigned long long p = a + lane; p < b; p += 32) {
      const uint32_t g = group_of_target[pat_targets[p]];
      if (atomicExch(&row[g], 0u) > 0u) atomicAdd(&hits[g], c);
    }
    __syncwarp();
  }
}

// Dense rows.  MODE 0: logl (fp64).  MODE 1: P = exp(logl - M) fp64 + M.  MODE 2: same in fp32.
// MODE 3: raw hit counts as uint32 (parity export; row stride K_all, all groups, pos = identity).
//
// LUT[g][0] = log(zero_inflation) = l0 for every group and almost every (class, group) pair has no hit, so a
// row is written as a constant with 16-byte stores and the few groups the class does hit are patched in by
// walking its pattern again: O(K / lanes) wide stores + O(pattern length) table lookups per class.  atomicExch
// hands each distinct group to one lane and leaves the counter row clean for the next class.
// M only has to be a shift that keeps P in range: max(l0, values of the hit groups) >= the true row maximum.
template <typename OT> struct Vec16;
template <> struct Vec16<double> { using type = double2; static constexpr int N = 2; static __device__ double2 splat(double v) { return make_double2(v, v); } };
template <> struct Vec16<float> { using type = float4; static constexpr int N = 4; static __device__ float4 splat(float v) { return make_float4(v, v, v, v); } };
template <> struct Vec16<uint32_t> { using type = uint4; static constexpr int N = 4; static __device__ uint4 splat(uint32_t v) { return make_uint4(v, v, v, v); } };

template <int MODE, typename OT>
__global__ void __launch_bounds__(LIK_NT)
lik_fill_kernel(const uint64_t *__restrict__ pat_ptr, const uint32_t *__restrict__ pat_targets,
                const uint32_t *__restrict__ group_of_target, const int *__restrict__ pos_of_group,
                const uint64_t *__restrict__ lut_off, const double *__restrict__ lut, unsigned long long N,
                int K_all, int K, int ld, int active_warps, double l0, OT *__restrict__ out, double *__restrict__ rowmax) {
  using V = Vec16<OT>;
  extern __shared__ unsigned s_rows[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp >= active_warps) return;   // (no block-wide barrier below: warps are independent)
  unsigned *row = s_rows + (size_t)warp * K_all;
  for (int k = lane; k < K_all; k += 32) row[k] = 0;
  __syncwarp();
  const int nvec = ld / V::N;         // ld is a multiple of the vector width (16-byte aligned rows)
  const unsigned long long n_warps = (unsigned long long)gridDim.x * active_warps;
  for (unsigned long long j = (unsigned long long)blockIdx.x * active_warps + warp; j < N; j += n_warps) {
    const unsigned long long a = pat_ptr[j], b = pat_ptr[j + 1];
    count_pattern(row, pat_targets, a, b, group_of_target, lane);
    OT *orow = out + j * (unsigned long long)ld;
    double m = l0;
    if (MODE == 1 || MODE == 2) {
      for (unsigned long long p = a + lane; p < b; p += 32) {
        const uint32_t g = group_of_target[pat_targets[p]];
        const int pos = pos_of_group[g];
        if (pos >= 0) m = fmax(m, lut[lut_off[pos] + row[g]]);
      }
      m = warp_max(m);
      if (lane == 0) rowmax[j] = m;
    }
    const OT fillv = MODE == 3 ? (OT)0 : (MODE == 0 ? (OT)l0 : (OT)exp(l0 - m));
    typename V::type *vrow = reinterpret_cast<typename V::type *>(orow);
    const typename V::type fv = V::splat(fillv);
    for (int i = lane; i < nvec; i += 32) vrow[i] = fv;
    if (MODE != 3 && ld > K) {          // padding columns hold 0
      __syncwarp();
      for (int k = K + lane; k < ld; k += 32) orow[k] = (OT)0;
    }
    __syncwarp();                       // the constant fill is ordered before the patches below
    for (unsigned long long p = a + lane; p < b; p += 32) {
      const uint32_t g = group_of_target[pat_targets[p]];
      const unsigned c = atomicExch(&row[g], 0u);
      if (c == 0) continue;             // another lane owns this group
      if (MODE == 3) { orow[g] = (OT)c; continue; }
      const int pos = pos_of_group[g];
      if (pos < 0) continue;            // pruned by --min-hits
      const double v = lut[lut_off[pos] + c];
      orow[pos] = (OT)(MODE == 0 ? v : exp(v - m));
    }
    __syncwarp();
  }
}

// ---- sparse storage -------------------------------------------------------------------------------------
// Pass 1: number of distinct kept groups each class hits.  Pass 2: the hits themselves, sorted by group inside a
// class so that every run produces the same arrays bit for bit.
__global__ void __launch_bounds__(LIK_NT)
sparse_count_kernel(const uint64_t *__restrict__ pat_ptr, const uint32_t *__restrict__ pat_targets,
                    const uint32_t *__restrict__ group_of_target, const int *__restrict__ pos_of_group,
                    unsigned long long N, int K_all, int active_warps, uint64_t *__restrict__ nz_count) {
  extern __shared__ unsigned s_rows[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp >= active_warps) return;
  unsigned *row = s_rows + (size_t)warp * K_all;
  for (int k = lane; k < K_all; k += 32) row[k] = 0;
  __syncwarp();
  const unsigned long long n_warps = (unsigned long long)gridDim.x * active_warps;
  for (unsigned long long j = (unsigned long long)blockIdx.x * active_warps + warp; j < N; j += n_warps) {
    const unsigned long long a = pat_ptr[j], b = pat_ptr[j + 1];
    count_pattern(row, pat_targets, a, b, group_of_target, lane);
    unsigned n = 0;
    for (unsigned long long p = a + lane; p < b; p += 32) {
      const uint32_t g = group_of_target[pat_targets[p]];
      if (atomicExch(&row[g], 0u) > 0u && pos_of_group[g] >= 0) ++n;
    }
    n = __reduce_add_sync(0xffffffffu, n);
    if (lane == 0) nz_count[j] = n;
    __syncwarp();
  }
}

__global__ void __launch_bounds__(LIK_NT)
sparse_fill_kernel(const uint64_t *__restrict__ pat_ptr, const uint32_t *__restrict__ pat_targets,
                   const uint32_t *__restrict__ group_of_target, const int *__restrict__ pos_of_group,
                   const uint64_t *__restrict__ lut_off, const double *__restrict__ lut, unsigned long long N,
                   int K_all, int active_warps, double l0, const uint64_t *__restrict__ nz_ptr,
                   uint32_t *__restrict__ nz_grp, double *__restrict__ nz_dP, double *__restrict__ nz_logl,
                   double *__restrict__ P0, double *__restrict__ rowmax) {
  extern __shared__ unsigned s_rows[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp >= active_warps) return;
  unsigned *row = s_rows + (size_t)warp * K_all;
  for (int k = lane; k < K_all; k += 32) row[k] = 0;
  __syncwarp();
  const unsigned long long n_warps = (unsigned long long)gridDim.x * active_warps;
  for (unsigned long long j = (unsigned long long)blockIdx.x * active_warps + warp; j < N; j += n_warps) {
    const unsigned long long a = pat_ptr[j], b = pat_ptr[j + 1];
    count_pattern(row, pat_targets, a, b, group_of_target, lane);
    double m = l0;
    for (unsigned long long p = a + lane; p < b; p += 32) {
      const uint32_t g = group_of_target[pat_targets[p]];
      const int pos = pos_of_group[g];
      if (pos >= 0) m = fmax(m, lut[lut_off[pos] + row[g]]);
    }
    m = warp_max(m);
    const double p0 = exp(l0 - m);
    if (lane == 0) { rowmax[j] = m; P0[j] = p0; }
    // emit the hits in arrival order ...
    const unsigned long long base = nz_ptr[j];
    const unsigned n_row = (unsigned)(nz_ptr[j + 1] - base);
    unsigned emitted = 0;
    for (unsigned long long p0i = a; p0i < b; p0i += 32) {
      const unsigned long long p = p0i + lane;
      bool mine = false;
      uint32_t pos_u = 0;
      double v = 0.0, lv = 0.0;
      if (p < b) {
        const uint32_t g = group_of_target[pat_targets[p]];
        const unsigned c = atomicExch(&row[g], 0u);
        const int pos = pos_of_group[g];
        // the sparse pass wants (class index inside its 32-class chunk, group) in one word: top 8 bits, low 24 bits
        if (c > 0u && pos >= 0) { mine = true; pos_u = (uint32_t)pos | ((uint32_t)(j & 31u) << 24); lv = lut[lut_off[pos] + c]; v = exp(lv - m) - p0; }
      }
      const unsigned ballot = __ballot_sync(0xffffffffu, mine);
      if (mine) {
        const unsigned slot = emitted + __popc(ballot & ((1u << lane) - 1u));
        nz_grp[base + slot] = pos_u;
        nz_dP[base + slot] = v;
        nz_logl[base + slot] = lv;
      }
      emitted += __popc(ballot);
    }
    __syncwarp();
    // ... then order them by group (distinct within a row): every entry finds its rank and moves there
    if (n_row <= 32) {
      uint32_t gi = 0; double vi = 0.0, li = 0.0; unsigned rank = 0;
      if ((unsigned)lane < n_row) {
        gi = nz_grp[base + lane]; vi = nz_dP[base + lane]; li = nz_logl[base + lane];
        for (unsigned q = 0; q < n_row; ++q) rank += nz_grp[base + q] < gi ? 1u : 0u;
      }
      __syncwarp();                                        // everything is read before anything is overwritten
      if ((unsigned)lane < n_row) { nz_grp[base + rank] = gi; nz_dP[base + rank] = vi; nz_logl[base + rank] = li; }
    } else if (lane == 0) {                                // a class hitting more than 32 groups: insertion sort, rare
      for (unsigned x = 1; x < n_row; ++x) {
        const uint32_t g = nz_grp[base + x]; const double v = nz_dP[base + x], l = nz_logl[base + x];
        unsigned y = x;
        while (y > 0 && nz_grp[base + y - 1] > g) {
          nz_grp[base + y] = nz_grp[base + y - 1]; nz_dP[base + y] = nz_dP[base + y - 1]; nz_logl[base + y] = nz_logl[base + y - 1]; --y;
        }
        nz_grp[base + y] = g; nz_dP[base + y] = v; nz_logl[base + y] = l;
      }
    }
    __syncwarp();
  }
}

__global__ void u64_to_double_kernel(const uint64_t *in, double *out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = (double)in[i];
}
__global__ void rebase_ptr_kernel(const uint64_t *in, uint64_t *out, size_t n, uint64_t base) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i] - base;
}

// in: rows x cols row-major  ->  out: cols x rows row-major with row stride ld_out (padding zeroed by the caller)
template <typename TI, typename TO>
__global__ void transpose_kernel(const TI *__restrict__ in, size_t rows, size_t cols, size_t ld_in,
                                 TO *__restrict__ out, size_t ld_out) {
  __shared__ TI tile[32][33];
  const size_t c0 = (size_t)blockIdx.x * 32, r0 = (size_t)blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const size_t r = r0 + i, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[i][threadIdx.x] = in[r * ld_in + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const size_t c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) out[c * ld_out + r] = (TO)tile[threadIdx.x][i];
  }
}

template <typename TI, typename TO>
void transpose(const TI *in, size_t rows, size_t cols, size_t ld_in, TO *out, size_t ld_out, cudaStream_t s) {
  if (rows == 0 || cols == 0) return;
  // gridDim.y is limited to 65535 blocks: walk the row dimension in slabs
  const size_t slab = (size_t)65535 * 32;
  for (size_t r0 = 0; r0 < rows; r0 += slab) {
    const size_t nr = std::min(slab, rows - r0);
    dim3 grid((unsigned)ceil_div(cols, 32), (unsigned)ceil_div(nr, 32));
    transpose_kernel<TI, TO><<<grid, dim3(32, 8), 0, s>>>(in + r0 * ld_in, nr, cols, ld_in, out + r0, ld_out);
    MSWB_LAUNCHED();
  }
}

} // namespace mswb

namespace {

// include/Likelihood.hpp:47-60, 92-107, 198-207 evaluated on the host with the same libm calls the
// reference makes; K'(S+1) values, negligible next to the matrix.  The table is ragged: group g' owns
// lut_off[g'] .. lut_off[g'] + size(g'), so one huge group does not inflate the others' rows.
double lbeta_h(double x, double y) { return std::lgamma(x) + std::lgamma(y) - std::lgamma(x + y); }
double ldbb_scaled_h(uint64_t k, uint64_t n, double alpha, double beta) {
  const double lbc = std::lgamma((double)n + 1.0) - std::lgamma((double)k + 1.0) - std::lgamma((double)(n - k) + 1.0);
  return lbc + lbeta_h((double)k + alpha, (double)(n - k) + beta) - lbeta_h((double)n + alpha, beta);
}

void build_lut(const std::vector<uint64_t> &sizes, double q, double e, double zi, std::vector<uint64_t> *off, std::vector<double> *lut) {
  off->assign(sizes.size() + 1, 0);
  for (size_t g = 0; g < sizes.size(); ++g) (*off)[g + 1] = (*off)[g] + sizes[g] + 1;
  lut->assign(off->back(), 0.0);
  const double l0 = std::log(zi), l1 = std::log1p(-zi);
  for (size_t g = 0; g < sizes.size(); ++g) {
    const double n = (double)sizes[g];
    const double mean = n * q;                       // update_bb_parameters, bb_constants = {q, e}
    const double phi = 1.0 / (n - mean + e);
    const double beta = phi * (n - mean);
    const double alpha = (mean * beta) / (n - mean);
    double *row = lut->data() + (*off)[g];
    row[0] = l0;
    for (uint64_t c = 1; c <= sizes[g]; ++c) row[c] = ldbb_scaled_h(c, sizes[g], alpha, beta) + l1;
  }
}

void device_exclusive_sum_u64(mswb_ctx *ctx, const uint64_t *in, uint64_t *out, size_t n) {
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int64_t)n, ctx->stream);
  mswb::DevBuf<unsigned char> tmp;
  tmp.alloc(bytes);
  MSWB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, in, out, (int64_t)n, ctx->stream));
  MSWB_CUDA(cudaStreamSynchronize(ctx->stream));
}

// Warps of a CTA that get a counter row: as many as fit in ~200 KB of shared memory.
int lik_active_warps(uint32_t K_all) {
  const size_t per_row = (size_t)K_all * sizeof(unsigned);
  MSWB_REQUIRE(per_row <= 200 * 1024, "too many groups for the shared-memory hit counters (max 51200)");
  return (int)std::min<size_t>(mswb::LIK_WARPS, (200 * 1024) / per_row);
}
size_t lik_smem_bytes(uint32_t K_all) { return (size_t)lik_active_warps(K_all) * K_all * sizeof(unsigned); }

template <class Kern> void prepare_fill_kernel(Kern kern, size_t smem) {
  MSWB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
}

int fill_grid(mswb_ctx *ctx, uint64_t N, uint32_t K_all) {
  const uint64_t want = ceil_div(N, (uint64_t)lik_active_warps(K_all));
  const uint64_t per_sm = std::max<size_t>(1, std::min<size_t>(4, (200 * 1024) / std::max<size_t>(1, lik_smem_bytes(K_all))));
  return (int)std::max<uint64_t>(1, std::min<uint64_t>(want, (uint64_t)ctx->n_sms * per_sm));
}

} // namespace

namespace mswb {

static bool is_large(const mswb_lik *L) {
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  return (size_t)L->N_pad * L->Kp * sizeof(double) > total_b / 4;
}

void lik_ensure_logl(mswb_lik *L) {
  if (L->logl.p) return;
  MSWB_REQUIRE(L->from_patterns, "the fp64 log-likelihood is not resident and cannot be rebuilt (no class patterns)");
  mswb_ctx *ctx = L->ctx;
  cudaStream_t s = ctx->stream;
  if (L->P64.p && is_large(L)) { MSWB_CUDA(cudaStreamSynchronize(s)); L->P64.release(); }
  L->logl.alloc((size_t)L->N * L->Kp);
  auto kern = lik_fill_kernel<0, double>;
  const size_t smem = lik_smem_bytes(L->K_all);
  prepare_fill_kernel(kern, smem);
  kern<<<fill_grid(ctx, L->N, L->K_all), LIK_NT, smem, s>>>(L->pat_ptr.p, L->pat_targets.p, L->group_of_target.p, L->pos_dev.p,
                                                            L->lut_off.p, L->lut.p, L->N, (int)L->K_all, (int)L->K, (int)L->Kp,
                                                            lik_active_warps(L->K_all), L->l0, L->logl.p, nullptr);
  MSWB_LAUNCHED();
}

template <typename ST> static void fill_linear(mswb_lik *L, DevBuf<ST> &P, uint32_t ld) {
  mswb_ctx *ctx = L->ctx;
  cudaStream_t s = ctx->stream;
  P.alloc((size_t)L->N_pad * ld);
  L->rowmax.alloc(L->N_pad);
  if (L->N_pad > L->N) MSWB_CUDA(cudaMemsetAsync(P.p + (size_t)L->N * ld, 0, (size_t)(L->N_pad - L->N) * ld * sizeof(ST), s));
  MSWB_CUDA(cudaMemsetAsync(L->rowmax.p, 0, L->rowmax.bytes(), s));
  auto kern = lik_fill_kernel<sizeof(ST) == 8 ? 1 : 2, ST>;
  const size_t smem = lik_smem_bytes(L->K_all);
  prepare_fill_kernel(kern, smem);
  kern<<<fill_grid(ctx, L->N, L->K_all), LIK_NT, smem, s>>>(L->pat_ptr.p, L->pat_targets.p, L->group_of_target.p, L->pos_dev.p,
                                                            L->lut_off.p, L->lut.p, L->N, (int)L->K_all, (int)L->K, (int)ld,
                                                            lik_active_warps(L->K_all), L->l0, P.p, L->rowmax.p);
  MSWB_LAUNCHED();
}

// logl (EC-major, ld) -> rowmax and P = exp(logl - rowmax), one warp per row (likelihoods given densely).
template <typename ST>
__global__ void to_linear_kernel(const double *__restrict__ logl, int ld, ST *__restrict__ P, int ldp,
                                 double *__restrict__ rowmax, unsigned long long N, int K) {
  const int lane = threadIdx.x & 31;
  const unsigned long long warp = (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) >> 5;
  const unsigned long long n_warps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
  for (unsigned long long row = warp; row < N; row += n_warps) {
    const double *lp = logl + row * (unsigned long long)ld;
    double m = -INFINITY;
    for (int k = lane; k < K; k += 32) m = fmax(m, lp[k]);
    m = warp_max(m);
    ST *pp = P + row * (unsigned long long)ldp;
    for (int k = lane; k < ldp; k += 32) pp[k] = k < K ? (ST)exp(lp[k] - m) : (ST)0;
    if (lane == 0) rowmax[row] = m;
  }
}

template <typename ST> static void dense_to_linear(mswb_lik *L, DevBuf<ST> &P, uint32_t ld) {
  mswb_ctx *ctx = L->ctx;
  cudaStream_t s = ctx->stream;
  MSWB_REQUIRE(L->logl.p, "likelihood holds neither the log-likelihood nor class patterns");
  P.alloc((size_t)L->N_pad * ld);
  L->rowmax.alloc(L->N_pad);
  if (L->N_pad > L->N) MSWB_CUDA(cudaMemsetAsync(P.p + (size_t)L->N * ld, 0, (size_t)(L->N_pad - L->N) * ld * sizeof(ST), s));
  MSWB_CUDA(cudaMemsetAsync(L->rowmax.p, 0, L->rowmax.bytes(), s));
  to_linear_kernel<ST><<<ctx->n_sms * 8, 256, 0, s>>>(L->logl.p, (int)L->Kp, P.p, (int)ld, L->rowmax.p, L->N, (int)L->K);
  MSWB_LAUNCHED();
}

void lik_ensure_sparse(mswb_lik *L) {
  if (L->nz_ptr.p) return;
  MSWB_REQUIRE(L->from_patterns, "sparse storage needs a likelihood built from class patterns");
  mswb_ctx *ctx = L->ctx;
  cudaStream_t s = ctx->stream;
  const size_t smem = lik_smem_bytes(L->K_all);
  const int grid = fill_grid(ctx, L->N, L->K_all), aw = lik_active_warps(L->K_all);
  DevBuf<uint64_t> cnt;
  cnt.alloc(L->N + 1);
  MSWB_CUDA(cudaMemsetAsync(cnt.p, 0, (L->N + 1) * sizeof(uint64_t), s));
  prepare_fill_kernel(sparse_count_kernel, smem);
  sparse_count_kernel<<<grid, LIK_NT, smem, s>>>(L->pat_ptr.p, L->pat_targets.p, L->group_of_target.p, L->pos_dev.p, L->N,
                                                 (int)L->K_all, aw, cnt.p);
  MSWB_LAUNCHED();
  L->nz_ptr.alloc(L->N + 1);
  device_exclusive_sum_u64(ctx, cnt.p, L->nz_ptr.p, L->N + 1);
  uint64_t nnz = 0;
  d2h(&nnz, L->nz_ptr.p + L->N, 1, s);
  MSWB_CUDA(cudaStreamSynchronize(s));
  L->nnz = nnz;
  L->nz_grp.alloc(nnz);
  L->nz_dP.alloc(nnz);
  L->nz_logl.alloc(nnz);
  L->P0.alloc(L->N_pad);
  L->rowmax.alloc(L->N_pad);
  MSWB_CUDA(cudaMemsetAsync(L->P0.p, 0, L->P0.bytes(), s));
  MSWB_CUDA(cudaMemsetAsync(L->rowmax.p, 0, L->rowmax.bytes(), s));
  prepare_fill_kernel(sparse_fill_kernel, smem);
  sparse_fill_kernel<<<grid, LIK_NT, smem, s>>>(L->pat_ptr.p, L->pat_targets.p, L->group_of_target.p, L->pos_dev.p, L->lut_off.p,
                                                L->lut.p, L->N, (int)L->K_all, aw, L->l0, L->nz_ptr.p, L->nz_grp.p, L->nz_dP.p,
                                                L->nz_logl.p, L->P0.p, L->rowmax.p);
  MSWB_LAUNCHED();
}

void lik_ensure_linear(mswb_lik *L) {
  if (L->storage == MSWB_STORE_F32) {
    if (L->P32.p) return;
    L->Kp32 = (uint32_t)round_up(L->K, 4);
    if (L->from_patterns) fill_linear<float>(L, L->P32, L->Kp32); else dense_to_linear<float>(L, L->P32, L->Kp32);
  } else {
    if (L->P64.p) return;
    if (L->from_patterns) {
      if (L->logl.p && is_large(L)) { MSWB_CUDA(cudaStreamSynchronize(L->ctx->stream)); L->logl.release(); }
      fill_linear<double>(L, L->P64, L->Kp);
    } else {
      dense_to_linear<double>(L, L->P64, L->Kp);
    }
  }
}

} // namespace mswb

extern "C" {

int mswb_lik_build(mswb_ctx *ctx, const mswb_aln *aln, const uint32_t *group_of_target, uint32_t n_groups,
                   const uint64_t *group_sizes, double q, double e, double zero_inflation, uint64_t min_hits,
                   int storage, mswb_lik **out) {
  return guarded([&] {
    MSWB_REQUIRE(ctx && aln && group_of_target && group_sizes && out, "NULL argument");
    MSWB_REQUIRE(aln->ctx == ctx, "alignment belongs to another context");
    MSWB_REQUIRE(storage == MSWB_STORE_F64 || storage == MSWB_STORE_F32 || storage == MSWB_STORE_SPARSE, "unknown storage");
    MSWB_REQUIRE(n_groups >= 1, "the grouping has no groups");
    MSWB_REQUIRE(aln->partitioned || aln->n_ecs > 0, "the alignment holds no equivalence class: no read aligned to any reference sequence");
    MSWB_REQUIRE(zero_inflation > 0.0 && zero_inflation < 1.0, "zero inflation must lie in (0, 1)");
    MSWB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const uint64_t T = aln->n_targets;
    // The hit count of a class in group g indexes row g of the ragged lookup table (0 .. size(g)): the sizes must be
    // the number of targets the indicators put in each group, exactly what the reference's Grouping holds
    // (include/Grouping.hpp:62-83); anything smaller would send the lookup into the next group's row.
    {
      std::vector<uint64_t> tally(n_groups, 0);
      for (uint64_t t = 0; t < T; ++t) {
        MSWB_REQUIRE(group_of_target[t] < n_groups, "group indicator out of range");
        ++tally[group_of_target[t]];
      }
      for (uint32_t g = 0; g < n_groups; ++g)
        MSWB_REQUIRE(group_sizes[g] == tally[g], "group_sizes[" + std::to_string(g) + "] = " + std::to_string(group_sizes[g]) +
                                                     " but the indicators put " + std::to_string(tally[g]) + " targets in that group");
    }
    MSWB_REQUIRE(storage != MSWB_STORE_SPARSE || n_groups < (1u << 24), "sparse storage packs the group in 24 bits");

    std::unique_ptr<mswb_lik> L(new mswb_lik);
    L->ctx = ctx;
    L->K_all = n_groups;
    L->storage = storage;
    uint64_t lo, hi;
    double n_aligned_total = (double)aln->n_aligned;
    if (aln->partitioned) {
      // every rank built the classes of its own hash range: the shard is the whole local table
      DevBuf<double> tmp;
      tmp.alloc(ctx->world + 1);
      std::vector<double> v(ctx->world + 1, 0.0);
      v[ctx->rank] = (double)aln->n_ecs;
      v[ctx->world] = (double)aln->n_aligned;
      h2d(tmp.p, v.data(), v.size(), s);
      ctx->allreduce_sum(tmp.p, v.size());
      d2h(v.data(), tmp.p, v.size(), s);
      MSWB_CUDA(cudaStreamSynchronize(s));
      ctx->peer_check();
      for (int r = 0; r < ctx->world; ++r) { if (r < ctx->rank) L->ec_begin += (uint64_t)v[r]; L->N_total += (uint64_t)v[r]; }
      n_aligned_total = v[ctx->world];
      lo = 0; hi = aln->n_ecs;
    } else {
      L->N_total = aln->n_ecs;
      MSWB_REQUIRE(mswb_shard_range(ctx, aln->n_ecs, &lo, &hi) == 0, mswb_last_error());
      L->ec_begin = lo;
    }
    L->N = hi - lo;
    L->N_pad = round_up(L->N, ROW_PAD);
    L->n_targets = T;
    L->from_patterns = true;
    L->l0 = std::log(zero_inflation);

    // shard of the pattern CSR, rebased
    uint64_t p_lo = 0, p_hi = 0;
    d2h(&p_lo, aln->pat_ptr.p + lo, 1, s);
    d2h(&p_hi, aln->pat_ptr.p + hi, 1, s);
    MSWB_CUDA(cudaStreamSynchronize(s));
    L->pat_ptr.alloc(L->N + 1);
    rebase_ptr_kernel<<<ctx->n_sms * 2, 256, 0, s>>>(aln->pat_ptr.p + lo, L->pat_ptr.p, L->N + 1, p_lo);
    MSWB_LAUNCHED();
    L->pat_targets.alloc(p_hi - p_lo);
    if (p_hi > p_lo)
      MSWB_CUDA(cudaMemcpyAsync(L->pat_targets.p, aln->pat_targets.p + p_lo, (p_hi - p_lo) * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    L->group_of_target.alloc(T);
    h2d(L->group_of_target.p, group_of_target, T, s);
    L->counts.alloc(L->N_pad);
    MSWB_CUDA(cudaMemsetAsync(L->counts.p, 0, L->counts.bytes(), s));
    u64_to_double_kernel<<<ctx->n_sms * 2, 256, 0, s>>>(aln->count.p + lo, L->counts.p, L->N);
    MSWB_LAUNCHED();
    L->sum_counts_total = n_aligned_total;   // every read with >= 1 hit sits in exactly one class

    const size_t smem = lik_smem_bytes(n_groups);
    const int grid = fill_grid(ctx, L->N, n_groups);
    const int aw = lik_active_warps(n_groups);

    // ---- --min-hits mask (include/Likelihood.hpp:141-171) ---------------------------------------
    L->mask.assign(n_groups, min_hits > 0 ? 0 : 1);
    if (min_hits > 0) {
      DevBuf<unsigned long long> hits_dev;
      hits_dev.alloc(n_groups);
      MSWB_CUDA(cudaMemsetAsync(hits_dev.p, 0, n_groups * sizeof(unsigned long long), s));
      prepare_fill_kernel(group_hits_kernel, smem);
      group_hits_kernel<<<grid, LIK_NT, smem, s>>>(L->pat_ptr.p, L->pat_targets.p, L->group_of_target.p, L->counts.p,
                                                   L->N, (int)n_groups, aw, hits_dev.p);
      MSWB_LAUNCHED();
      ctx->allreduce_sum_u64(hits_dev.p, n_groups);
      L->hits.resize(n_groups);
      static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "");
      d2h((unsigned long long *)L->hits.data(), hits_dev.p, n_groups, s);
      MSWB_CUDA(cudaStreamSynchronize(s));
      for (uint32_t g = 0; g < n_groups; ++g) L->mask[g] = L->hits[g] >= min_hits ? 1 : 0;
    }
    std::vector<uint64_t> kept_sizes;
    for (uint32_t g = 0; g < n_groups; ++g)
      if (L->mask[g]) { L->kept.push_back(g); kept_sizes.push_back(group_sizes[g]); }
    L->K = (uint32_t)L->kept.size();
    MSWB_REQUIRE(L->K >= 1, "--min-hits removed every group");
    L->Kp = (uint32_t)round_up(L->K, 2);
    L->kept_dev.alloc(L->K);
    h2d(L->kept_dev.p, L->kept.data(), L->K, s);
    std::vector<int> pos(n_groups, -1);
    for (uint32_t k = 0; k < L->K; ++k) pos[L->kept[k]] = (int)k;
    L->pos_dev.alloc(n_groups);
    h2d(L->pos_dev.p, pos.data(), n_groups, s);

    // ---- lookup table (include/Likelihood.hpp:92-107) --------------------------------------------
    std::vector<uint64_t> off;
    std::vector<double> lut;
    build_lut(kept_sizes, q, e, zero_inflation, &off, &lut);
    L->lut_off.alloc(off.size());
    L->lut.alloc(lut.size());
    h2d(L->lut_off.p, off.data(), off.size(), s);
    h2d(L->lut.p, lut.data(), lut.size(), s);

    // ---- dense gather (include/Likelihood.hpp:176-185), EC-major ---------------------------------
    // fp32 storage exists for matrices that do not fit as fp64: it goes straight to the linear form.
    if (storage == MSWB_STORE_F64) lik_ensure_logl(L.get());
    else if (storage == MSWB_STORE_F32) lik_ensure_linear(L.get());
    else lik_ensure_sparse(L.get());
    MSWB_CUDA(cudaStreamSynchronize(s));   // host vectors above go out of scope
    *out = L.release();
  });
}

int mswb_lik_from_dense(mswb_ctx *ctx, const double *logl, uint32_t n_groups, uint64_t n_ecs_local,
                        const double *log_counts, int storage, mswb_lik **out) {
  return guarded([&] {
    MSWB_REQUIRE(ctx && logl && log_counts && out, "NULL argument");
    MSWB_REQUIRE(storage == MSWB_STORE_F64 || storage == MSWB_STORE_F32, "unknown storage");
    MSWB_REQUIRE(n_groups >= 1, "the matrix has no groups");
    MSWB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    std::unique_ptr<mswb_lik> L(new mswb_lik);
    L->ctx = ctx;
    L->K_all = L->K = n_groups;
    L->Kp = (uint32_t)round_up(n_groups, 2);
    L->N = n_ecs_local;
    L->N_pad = round_up(n_ecs_local, ROW_PAD);
    L->storage = storage;
    L->mask.assign(n_groups, 1);
    L->kept.resize(n_groups);
    for (uint32_t g = 0; g < n_groups; ++g) L->kept[g] = g;

    // global class bookkeeping over the ranks
    DevBuf<double> tmp;
    tmp.alloc(ctx->world + 1);
    std::vector<double> sizes(ctx->world, 0.0);
    sizes[ctx->rank] = (double)n_ecs_local;
    h2d(tmp.p, sizes.data(), ctx->world, s);
    ctx->allreduce_sum(tmp.p, ctx->world);
    d2h(sizes.data(), tmp.p, ctx->world, s);
    MSWB_CUDA(cudaStreamSynchronize(s));
    ctx->peer_check();
    for (int r = 0; r < ctx->world; ++r) { if (r < ctx->rank) L->ec_begin += (uint64_t)sizes[r]; L->N_total += (uint64_t)sizes[r]; }

    // K x N group-major on the host -> N x Kp EC-major on the device
    const size_t n_el = (size_t)n_groups * n_ecs_local;
    DevBuf<double> staging;
    staging.alloc(n_el);
    h2d(staging.p, logl, n_el, s);
    L->logl.alloc((size_t)L->N * L->Kp);
    MSWB_CUDA(cudaMemsetAsync(L->logl.p, 0, L->logl.bytes(), s));
    transpose<double, double>(staging.p, n_groups, n_ecs_local, n_ecs_local, L->logl.p, L->Kp, s);

    // counts = exp(log_counts) rounded back to the integer they came from when they are one
    std::vector<double> c(n_ecs_local);
    long double total = 0.0L;
    for (uint64_t j = 0; j < n_ecs_local; ++j) {
      double v = std::exp(log_counts[j]);
      const double r = std::nearbyint(v);
      if (std::fabs(v - r) < 1e-9 * std::max(1.0, r)) v = r;
      c[j] = v;
      total += v;
    }
    L->counts.alloc(L->N_pad);
    MSWB_CUDA(cudaMemsetAsync(L->counts.p, 0, L->counts.bytes(), s));
    h2d(L->counts.p, c.data(), n_ecs_local, s);
    double tot = (double)total;
    h2d(tmp.p, &tot, 1, s);
    ctx->allreduce_sum(tmp.p, 1);
    d2h(&L->sum_counts_total, tmp.p, 1, s);
    MSWB_CUDA(cudaStreamSynchronize(s));
    ctx->peer_check();
    *out = L.release();
  });
}

int mswb_lik_info(const mswb_lik *lik, uint32_t *n_groups_all, uint32_t *n_groups_kept, uint64_t *n_ecs_local,
                  uint64_t *ec_begin, uint64_t *n_ecs_total) {
  return guarded([&] {
    MSWB_REQUIRE(lik, "lik is NULL");
    if (n_groups_all) *n_groups_all = lik->K_all;
    if (n_groups_kept) *n_groups_kept = lik->K;
    if (n_ecs_local) *n_ecs_local = lik->N;
    if (ec_begin) *ec_begin = lik->ec_begin;
    if (n_ecs_total) *n_ecs_total = lik->N_total;
  });
}

int mswb_lik_mask(const mswb_lik *lik, uint8_t *mask, uint64_t *hits) {
  return guarded([&] {
    MSWB_REQUIRE(lik, "lik is NULL");
    if (mask) std::memcpy(mask, lik->mask.data(), lik->K_all);
    if (hits) {
      MSWB_REQUIRE(!lik->hits.empty(), "no --min-hits tallies were computed (min_hits == 0)");
      std::memcpy(hits, lik->hits.data(), lik->K_all * sizeof(uint64_t));
    }
  });
}

int mswb_lik_export_hit_counts(const mswb_lik *lik, uint32_t *out) {
  return guarded([&] {
    MSWB_REQUIRE(lik && out, "NULL argument");
    MSWB_REQUIRE(lik->from_patterns, "this likelihood was not built from class patterns");
    mswb_ctx *ctx = lik->ctx;
    MSWB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const size_t n_el = (size_t)lik->N * lik->K_all;
    if (n_el == 0) return;
    const size_t ld = round_up(lik->K_all, 4);          // 16-byte aligned rows for the vector fill
    DevBuf<uint32_t> ecmajor, gmajor;
    ecmajor.alloc((size_t)lik->N * ld);
    gmajor.alloc(n_el);
    auto kern = lik_fill_kernel<3, uint32_t>;
    const size_t smem = lik_smem_bytes(lik->K_all);
    prepare_fill_kernel(kern, smem);
    kern<<<fill_grid(ctx, lik->N, lik->K_all), LIK_NT, smem, s>>>(lik->pat_ptr.p, lik->pat_targets.p, lik->group_of_target.p, nullptr,
                                                                  nullptr, nullptr, lik->N, (int)lik->K_all, (int)lik->K_all, (int)ld,
                                                                  lik_active_warps(lik->K_all), 0.0, ecmajor.p, nullptr);
    MSWB_LAUNCHED();
    transpose<uint32_t, uint32_t>(ecmajor.p, lik->N, lik->K_all, ld, gmajor.p, lik->N, s);
    d2h(out, gmajor.p, n_el, s);
    MSWB_CUDA(cudaStreamSynchronize(s));
  });
}

int mswb_lik_export_logl(const mswb_lik *lik, double *out) {
  return guarded([&] {
    MSWB_REQUIRE(lik && out, "NULL argument");
    MSWB_REQUIRE(lik->storage == MSWB_STORE_F64, "the fp64 log-likelihood is not kept in fp32 storage (only the linear form is)");
    lik_ensure_logl(const_cast<mswb_lik *>(lik));
    mswb_ctx *ctx = lik->ctx;
    MSWB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const size_t n_el = (size_t)lik->N * lik->K;
    if (n_el == 0) return;
    DevBuf<double> gmajor;
    gmajor.alloc(n_el);
    transpose<double, double>(lik->logl.p, lik->N, lik->K, lik->Kp, gmajor.p, lik->N, s);
    d2h(out, gmajor.p, n_el, s);
    MSWB_CUDA(cudaStreamSynchronize(s));
  });
}

void mswb_lik_destroy(mswb_lik *lik) {
  if (!lik) return;
  cudaSetDevice(lik->ctx->device);
  delete lik;
}

} // extern "C"
