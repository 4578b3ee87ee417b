// assign.cu — thresholded class -> group assignment: the device side of the --bin-reads hand-off.
//
// The reference passes the read ids of every class (Alignment::get_aligned_reads,
// include/mSWEEP_alignment.hpp:241; Sample::store_aligned_reads, include/Sample.hpp:89-105), the
// abundances and the K x N log-posteriors to mGEMS::BinFromMatrix (src/mSWEEP.cpp:437-469), which puts a
// class — all of its reads — into the bin of group k when the class's posterior for k reaches the
// group's threshold (by default the group's abundance).  mGEMS v1.3.3 is not vendored in the reference
// tree: the rule is restated from its published description, PARITY UNPINNED.
//
// Here the K x N matrix never leaves the device: posteriors are recomputed tile by tile, every (class,
// group) pair that passes emits (group << 32 | read id) for the class's reads, and one radix sort puts the
// bins in group order with ascending read ids inside.
#include "handles.cuh"

#include <cub/cub.cuh>
#include <memory>

using namespace mswb;

namespace mswb {

// One thread per class of the tile: how many groups take it, and the reads each group gains.
__global__ void assign_count_kernel(const double *__restrict__ tile, unsigned long long n_rows, int K,
                                    const double *__restrict__ log_thr, const uint64_t *__restrict__ ec_count,
                                    uint64_t *__restrict__ pairs, unsigned long long *__restrict__ bin_cnt) {
  for (unsigned long long j = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; j < n_rows;
       j += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long c = ec_count[j];
    unsigned long long m = 0;
    for (int k = 0; k < K; ++k)
      if (tile[(size_t)k * n_rows + j] >= log_thr[k]) { ++m; atomicAdd(&bin_cnt[k], c); }
    pairs[j] = m * c;
  }
}

// One warp per class: for each passing group (ascending), one key per member read.
__global__ void assign_emit_kernel(const double *__restrict__ tile, unsigned long long n_rows, int K,
                                   const double *__restrict__ log_thr, const uint64_t *__restrict__ read_ptr,
                                   const uint32_t *__restrict__ read_ids, const uint64_t *__restrict__ pair_off,
                                   uint64_t *__restrict__ keys) {
  const int lane = threadIdx.x & 31;
  const unsigned long long warp = (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) >> 5;
  const unsigned long long n_warps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
  for (unsigned long long j = warp; j < n_rows; j += n_warps) {
    const unsigned long long r0 = read_ptr[j], c = read_ptr[j + 1] - r0;
    unsigned long long out = pair_off[j];
    for (int k0 = 0; k0 < K; k0 += 32) {
      const int k = k0 + lane;
      const bool pass = k < K && tile[(size_t)k * n_rows + j] >= log_thr[k];
      unsigned hits = __ballot_sync(0xffffffffu, pass);
      while (hits) {
        const int b = __ffs(hits) - 1;
        hits &= hits - 1;
        const uint64_t hi = (uint64_t)(k0 + b) << 32;
        for (unsigned long long r = lane; r < c; r += 32) keys[out + r] = hi | read_ids[r0 + r];
        out += c;
      }
    }
  }
}

__global__ void low_words_kernel(const uint64_t *__restrict__ keys, unsigned long long n, uint32_t *__restrict__ out) {
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x) out[i] = (uint32_t)keys[i];
}

} // namespace mswb

extern "C" {

int mswb_vi_assign(mswb_ctx *ctx, mswb_lik *lik, const mswb_aln *aln, const double *log_threshold, uint64_t *bin_ptr) {
  return guarded([&] {
    MSWB_REQUIRE(ctx && lik && aln && log_threshold && bin_ptr, "NULL argument");
    MSWB_REQUIRE(aln->partitioned ? aln->n_ecs == lik->N : aln->n_ecs == lik->N_total,
                 "the alignment does not belong to this likelihood (class counts differ)");
    MSWB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const uint64_t N = lik->N, first = aln->partitioned ? 0 : lik->ec_begin;
    const int K = (int)lik->K;
    const int grid = ctx->n_sms * 8;
    lik->assign_reads.release();
    lik->assign_total = 0;

    DevBuf<double> thr, tile;
    DevBuf<unsigned long long> bin_cnt;
    DevBuf<uint64_t> pairs, pair_off;
    thr.alloc(K); bin_cnt.alloc(K); pairs.alloc(N + 1); pair_off.alloc(N + 1);
    h2d(thr.p, log_threshold, K, s);
    MSWB_CUDA(cudaMemsetAsync(bin_cnt.p, 0, (size_t)K * sizeof(unsigned long long), s));
    MSWB_CUDA(cudaMemsetAsync(pairs.p, 0, (N + 1) * sizeof(uint64_t), s));
    // tiles of at most 256 MB of posteriors
    const uint64_t tile_rows = std::max<uint64_t>(1, std::min<uint64_t>(N, (256ull << 20) / ((uint64_t)K * sizeof(double))));
    tile.alloc((size_t)tile_rows * K);
    for (uint64_t b = 0; b < N; b += tile_rows) {
      const uint64_t n = std::min(tile_rows, N - b);
      posterior_tile_dev(ctx, lik, b, n, tile.p);
      assign_count_kernel<<<grid, 256, 0, s>>>(tile.p, n, K, thr.p, aln->count.p + first + b, pairs.p + b, bin_cnt.p);
      MSWB_LAUNCHED();
    }
    size_t tmp_bytes = 0;
    DevBuf<unsigned char> tmp;
    MSWB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, pairs.p, pair_off.p, (int64_t)N + 1, s));
    tmp.alloc(tmp_bytes);
    MSWB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, pairs.p, pair_off.p, (int64_t)N + 1, s));
    uint64_t total = 0;
    std::vector<unsigned long long> cnt(K);
    d2h(&total, pair_off.p + N, 1, s);
    d2h(cnt.data(), bin_cnt.p, K, s);
    MSWB_CUDA(cudaStreamSynchronize(s));
    bin_ptr[0] = 0;
    for (int k = 0; k < K; ++k) bin_ptr[k + 1] = bin_ptr[k] + cnt[k];
    MSWB_REQUIRE(bin_ptr[K] == total, "internal error: bin sizes do not add up");
    if (total == 0) return;

    DevBuf<uint64_t> keys, keys_sorted;
    keys.alloc(total); keys_sorted.alloc(total);
    for (uint64_t b = 0; b < N; b += tile_rows) {
      const uint64_t n = std::min(tile_rows, N - b);
      posterior_tile_dev(ctx, lik, b, n, tile.p);
      assign_emit_kernel<<<grid, 256, 0, s>>>(tile.p, n, K, thr.p, aln->read_ptr.p + first + b, aln->read_ids.p, pair_off.p + b, keys.p);
      MSWB_LAUNCHED();
    }
    int group_bits = 1;
    while ((1ull << group_bits) < (uint64_t)K) ++group_bits;
    size_t sort_bytes = 0;
    MSWB_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, sort_bytes, keys.p, keys_sorted.p, (int64_t)total, 0, 32 + group_bits, s));
    DevBuf<unsigned char> sort_tmp;
    sort_tmp.alloc(sort_bytes);
    MSWB_CUDA(cub::DeviceRadixSort::SortKeys(sort_tmp.p, sort_bytes, keys.p, keys_sorted.p, (int64_t)total, 0, 32 + group_bits, s));
    lik->assign_reads.alloc(total);
    low_words_kernel<<<grid, 256, 0, s>>>(keys_sorted.p, total, lik->assign_reads.p);
    MSWB_LAUNCHED();
    MSWB_CUDA(cudaStreamSynchronize(s));
    lik->assign_total = total;
  });
}

int mswb_vi_assign_fetch(mswb_lik *lik, uint32_t *read_ids) {
  return guarded([&] {
    MSWB_REQUIRE(lik, "lik is NULL");
    if (lik->assign_total) {
      MSWB_REQUIRE(read_ids, "read_ids is NULL");
      MSWB_CUDA(cudaSetDevice(lik->ctx->device));
      d2h(read_ids, lik->assign_reads.p, lik->assign_total, lik->ctx->stream);
      MSWB_CUDA(cudaStreamSynchronize(lik->ctx->stream));
    }
    lik->assign_reads.release();
    lik->assign_total = 0;
  });
}

} // extern "C"
