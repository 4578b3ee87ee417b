// ec_build.cu — equivalence classes on the device (Alignment::collapse,
// include/mSWEEP_alignment.hpp:137-215 of the reference).
//
// The reference hashes every read's hit pattern, groups reads BY HASH in per-thread unordered_maps,
// merges them into a std::map (ascending hash) and keeps the pattern of the first read of each class.
// Here: one fold per read over its CSR row, a stable radix sort of (hash, read id), a run-length
// encode.  Classes come out in ascending unsigned hash order with ascending read ids inside, exactly
// the std::map iteration order, and colliding patterns merge exactly as they do in the reference.
#include "handles.cuh"

#include <cub/cub.cuh>
#include <memory>

using namespace mswb;

namespace mswb {

// hash ^= j + 0x517cc1b727220a95 + (hash << 6) + (hash >> 2) over the ascending set bits j
// (include/mSWEEP_alignment.hpp:150-155).  Also validates the input contract (ascending, < T).
__global__ void read_hash_kernel(const uint64_t *__restrict__ row_ptr, const uint32_t *__restrict__ targets,
                                 unsigned long long R, unsigned long long T, uint64_t *__restrict__ hash,
                                 unsigned char *__restrict__ aligned, int *__restrict__ bad) {
  for (unsigned long long r = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; r < R;
       r += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long a = row_ptr[r], b = row_ptr[r + 1];
    uint64_t h = 0;
    uint32_t prev = 0;
    for (unsigned long long p = a; p < b; ++p) {
      const uint32_t j = targets[p];
      if ((p > a && j <= prev) || j >= T) *bad = 1;
      prev = j;
      h ^= (uint64_t)j + 0x517cc1b727220a95ULL + (h << 6) + (h >> 2);
    }
    hash[r] = h;
    aligned[r] = b > a ? 1 : 0;
  }
}

__global__ void iota_kernel(uint32_t *out, unsigned long long n) {
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x) out[i] = (uint32_t)i;
}

// head[i] = 1 where a new hash value starts in the sorted key array
__global__ void head_flags_kernel(const uint64_t *__restrict__ keys, unsigned long long n, uint32_t *__restrict__ head) {
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x) head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}

// class e starts at sorted position i: record hash, first (= smallest) read id, start offset, pattern length
__global__ void scatter_heads_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ ids,
                                     const uint32_t *__restrict__ head, const uint32_t *__restrict__ ec_of,
                                     unsigned long long n, const uint64_t *__restrict__ row_ptr,
                                     uint64_t *__restrict__ ec_hash, uint32_t *__restrict__ rep_read,
                                     uint64_t *__restrict__ read_ptr, uint64_t *__restrict__ pat_len) {
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    if (!head[i]) continue;
    const uint32_t e = ec_of[i];
    const uint32_t r = ids[i];
    ec_hash[e] = keys[i];
    rep_read[e] = r;
    read_ptr[e] = i;
    pat_len[e] = row_ptr[r + 1] - row_ptr[r];
  }
}

__global__ void class_counts_kernel(const uint64_t *__restrict__ read_ptr, unsigned long long n_ecs, uint64_t *__restrict__ count) {
  for (unsigned long long e = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; e < n_ecs;
       e += (unsigned long long)gridDim.x * blockDim.x) count[e] = read_ptr[e + 1] - read_ptr[e];
}

// one warp copies the representative read's row into the class pattern CSR
__global__ void gather_patterns_kernel(const uint64_t *__restrict__ row_ptr, const uint32_t *__restrict__ targets,
                                       const uint32_t *__restrict__ rep_read, const uint64_t *__restrict__ pat_ptr,
                                       unsigned long long n_ecs, uint32_t *__restrict__ pat_targets) {
  const int lane = threadIdx.x & 31;
  const unsigned long long warp = (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) >> 5;
  const unsigned long long n_warps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
  for (unsigned long long e = warp; e < n_ecs; e += n_warps) {
    const unsigned long long src = row_ptr[rep_read[e]], dst = pat_ptr[e], len = pat_ptr[e + 1] - dst;
    for (unsigned long long p = lane; p < len; p += 32) pat_targets[dst + p] = targets[src + p];
  }
}

} // namespace mswb

static int ec_build_impl(mswb_ctx *ctx, uint64_t n_reads, uint64_t n_targets, const uint64_t *row_ptr,
                         const uint32_t *targets, bool partitioned, mswb_aln **out) {
  return guarded([&] {
    MSWB_REQUIRE(ctx && row_ptr && out, "NULL argument");
    MSWB_REQUIRE(n_reads <= 0xFFFFFFFFull, "more than 2^32 reads (the reference stores read ids as uint32_t, mSWEEP_alignment.hpp:45)");
    MSWB_REQUIRE(n_targets >= 1 && n_targets <= 0xFFFFFFFFull, "bad number of targets");
    MSWB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const uint64_t R = n_reads, nnz = row_ptr[R];
    MSWB_REQUIRE(nnz == 0 || targets, "targets is NULL");
    const int grid = ctx->n_sms * 8;

    std::unique_ptr<mswb_aln> A(new mswb_aln);
    A->ctx = ctx; A->n_reads = R; A->n_targets = n_targets; A->partitioned = partitioned;

    DevBuf<uint64_t> d_row_ptr, d_hash, d_keys_in, d_keys_out, d_pat_len;
    DevBuf<uint32_t> d_targets, d_ids, d_ids_in, d_ids_out, d_head, d_ec_of;
    DevBuf<unsigned char> d_aligned, d_tmp;
    DevBuf<int> d_flag;
    DevBuf<unsigned long long> d_num;
    d_row_ptr.alloc(R + 1); d_targets.alloc(nnz);
    h2d(d_row_ptr.p, row_ptr, R + 1, s);
    h2d(d_targets.p, targets, nnz, s);
    d_hash.alloc(R); d_aligned.alloc(R); d_ids.alloc(R); d_flag.alloc(1); d_num.alloc(1);
    MSWB_CUDA(cudaMemsetAsync(d_flag.p, 0, sizeof(int), s));

    // K1: per-read hash
    if (R) {
      read_hash_kernel<<<grid, 256, 0, s>>>(d_row_ptr.p, d_targets.p, R, n_targets, d_hash.p, d_aligned.p, d_flag.p);
      MSWB_LAUNCHED();
      iota_kernel<<<grid, 256, 0, s>>>(d_ids.p, R);
      MSWB_LAUNCHED();
    }
    int bad = 0;
    d2h(&bad, d_flag.p, 1, s);

    // keep the reads with at least one hit (mSWEEP_alignment.hpp:149), preserving read order
    d_keys_in.alloc(R); d_ids_in.alloc(R);
    size_t tmp_bytes = 0, need = 0;
    cub::DeviceSelect::Flagged(nullptr, need, d_hash.p, d_aligned.p, d_keys_in.p, d_num.p, (int64_t)R, s); tmp_bytes = std::max(tmp_bytes, need);
    cub::DeviceSelect::Flagged(nullptr, need, d_ids.p, d_aligned.p, d_ids_in.p, d_num.p, (int64_t)R, s); tmp_bytes = std::max(tmp_bytes, need);
    cub::DeviceRadixSort::SortPairs(nullptr, need, d_keys_in.p, d_keys_in.p, d_ids_in.p, d_ids_in.p, (int64_t)R, 0, 64, s); tmp_bytes = std::max(tmp_bytes, need);
    cub::DeviceScan::ExclusiveSum(nullptr, need, d_ids.p, d_ids.p, (int64_t)R, s); tmp_bytes = std::max(tmp_bytes, need);
    cub::DeviceScan::ExclusiveSum(nullptr, need, d_hash.p, d_hash.p, (int64_t)R + 1, s); tmp_bytes = std::max(tmp_bytes, need);
    d_tmp.alloc(tmp_bytes);
    MSWB_CUDA(cub::DeviceSelect::Flagged(d_tmp.p, tmp_bytes, d_hash.p, d_aligned.p, d_keys_in.p, d_num.p, (int64_t)R, s));
    MSWB_CUDA(cub::DeviceSelect::Flagged(d_tmp.p, tmp_bytes, d_ids.p, d_aligned.p, d_ids_in.p, d_num.p, (int64_t)R, s));
    unsigned long long n_al = 0;
    d2h(&n_al, d_num.p, 1, s);
    MSWB_CUDA(cudaStreamSynchronize(s));
    MSWB_REQUIRE(!bad, "pseudoalignment rows must hold strictly ascending target ids below n_targets");
    A->n_aligned = n_al;

    // K2: stable LSD radix sort by hash (read ids stay ascending inside equal hashes), then run-length encode.
    // Everything up to the pattern offsets is enqueued against capacity-n_al scratch, so that the class count and the
    // pattern size come back in ONE round trip (two host synchronisations per build: this one and the one above).
    d_keys_out.alloc(n_al); d_ids_out.alloc(n_al); d_head.alloc(n_al); d_ec_of.alloc(n_al);
    DevBuf<uint64_t> t_hash, t_read_ptr, t_pat_ptr;
    DevBuf<uint32_t> t_rep;
    t_hash.alloc(n_al); t_rep.alloc(n_al); t_read_ptr.alloc(n_al + 1); t_pat_ptr.alloc(n_al + 1);
    d_pat_len.alloc(n_al + 1);
    MSWB_CUDA(cudaMemsetAsync(d_pat_len.p, 0, (n_al + 1) * sizeof(uint64_t), s));
    uint64_t n_ecs = 0, pat_nnz = 0;
    if (n_al) {
      MSWB_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp.p, tmp_bytes, d_keys_in.p, d_keys_out.p, d_ids_in.p, d_ids_out.p, (int64_t)n_al, 0, 64, s));
      head_flags_kernel<<<grid, 256, 0, s>>>(d_keys_out.p, n_al, d_head.p);
      MSWB_LAUNCHED();
      MSWB_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp.p, tmp_bytes, d_head.p, d_ec_of.p, (int64_t)n_al, s));
      scatter_heads_kernel<<<grid, 256, 0, s>>>(d_keys_out.p, d_ids_out.p, d_head.p, d_ec_of.p, n_al, d_row_ptr.p,
                                                t_hash.p, t_rep.p, t_read_ptr.p, d_pat_len.p);
      MSWB_LAUNCHED();
      // pattern lengths sit at [0, n_ecs), zeros behind: the scan over the whole scratch ends in the pattern size
      MSWB_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp.p, tmp_bytes, d_pat_len.p, t_pat_ptr.p, (int64_t)n_al + 1, s));
      uint32_t last_idx = 0, last_head = 0;
      d2h(&last_idx, d_ec_of.p + (n_al - 1), 1, s);
      d2h(&last_head, d_head.p + (n_al - 1), 1, s);
      d2h(&pat_nnz, t_pat_ptr.p + n_al, 1, s);
      MSWB_CUDA(cudaStreamSynchronize(s));
      n_ecs = (uint64_t)last_idx + last_head;
    }
    A->n_ecs = n_ecs;
    A->pat_nnz = pat_nnz;
    A->hash.alloc(n_ecs); A->count.alloc(n_ecs); A->rep_read.alloc(n_ecs);
    A->read_ptr.alloc(n_ecs + 1); A->pat_ptr.alloc(n_ecs + 1);
    A->pat_targets.alloc(pat_nnz);
    if (n_ecs) {
      MSWB_CUDA(cudaMemcpyAsync(A->hash.p, t_hash.p, n_ecs * sizeof(uint64_t), cudaMemcpyDeviceToDevice, s));
      MSWB_CUDA(cudaMemcpyAsync(A->rep_read.p, t_rep.p, n_ecs * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
      MSWB_CUDA(cudaMemcpyAsync(A->read_ptr.p, t_read_ptr.p, n_ecs * sizeof(uint64_t), cudaMemcpyDeviceToDevice, s));
      MSWB_CUDA(cudaMemcpyAsync(A->pat_ptr.p, t_pat_ptr.p, (n_ecs + 1) * sizeof(uint64_t), cudaMemcpyDeviceToDevice, s));
    } else {
      MSWB_CUDA(cudaMemsetAsync(A->pat_ptr.p, 0, sizeof(uint64_t), s));
    }
    h2d(A->read_ptr.p + n_ecs, (const uint64_t *)&A->n_aligned, 1, s);
    if (n_ecs) {
      class_counts_kernel<<<grid, 256, 0, s>>>(A->read_ptr.p, n_ecs, A->count.p);
      MSWB_LAUNCHED();
      gather_patterns_kernel<<<grid, 256, 0, s>>>(d_row_ptr.p, d_targets.p, A->rep_read.p, A->pat_ptr.p, n_ecs, A->pat_targets.p);
      MSWB_LAUNCHED();
    }
    A->read_ids = std::move(d_ids_out);
    // (no further synchronisation: everything downstream runs on the same stream, and the scratch buffers above are
    // released through cudaFree / the block cache, both of which wait for the device)
    *out = A.release();
  });
}

extern "C" {

int mswb_ec_build(mswb_ctx *ctx, uint64_t n_reads, uint64_t n_targets, const uint64_t *row_ptr,
                  const uint32_t *targets, mswb_aln **out) {
  return ec_build_impl(ctx, n_reads, n_targets, row_ptr, targets, false, out);
}
int mswb_ec_build_partitioned(mswb_ctx *ctx, uint64_t n_reads_local, uint64_t n_targets, const uint64_t *row_ptr,
                              const uint32_t *targets, mswb_aln **out) {
  return ec_build_impl(ctx, n_reads_local, n_targets, row_ptr, targets, true, out);
}
uint64_t mswb_pattern_hash(const uint32_t *targets, uint64_t n) {
  uint64_t h = 0;
  for (uint64_t a = 0; a < n; ++a) h ^= (uint64_t)targets[a] + 0x517cc1b727220a95ULL + (h << 6) + (h >> 2);
  return h;
}

int mswb_ec_info(const mswb_aln *aln, uint64_t *n_ecs, uint64_t *n_reads, uint64_t *n_aligned, uint64_t *pattern_nnz) {
  return guarded([&] {
    MSWB_REQUIRE(aln, "aln is NULL");
    if (n_ecs) *n_ecs = aln->n_ecs;
    if (n_reads) *n_reads = aln->n_reads;
    if (n_aligned) *n_aligned = aln->n_aligned;
    if (pattern_nnz) *pattern_nnz = aln->pat_nnz;
  });
}

int mswb_ec_export(const mswb_aln *aln, uint64_t *hash, uint64_t *count, uint32_t *rep_read, uint64_t *pat_ptr,
                   uint32_t *pat_targets, uint64_t *read_ptr, uint32_t *read_ids) {
  return guarded([&] {
    MSWB_REQUIRE(aln, "aln is NULL");
    MSWB_CUDA(cudaSetDevice(aln->ctx->device));
    cudaStream_t s = aln->ctx->stream;
    const uint64_t n = aln->n_ecs;
    if (hash) d2h(hash, aln->hash.p, n, s);
    if (count) d2h(count, aln->count.p, n, s);
    if (rep_read) d2h(rep_read, aln->rep_read.p, n, s);
    if (pat_ptr) d2h(pat_ptr, aln->pat_ptr.p, n + 1, s);
    if (pat_targets) d2h(pat_targets, aln->pat_targets.p, aln->pat_nnz, s);
    if (read_ptr) d2h(read_ptr, aln->read_ptr.p, n + 1, s);
    if (read_ids) d2h(read_ids, aln->read_ids.p, aln->n_aligned, s);
    MSWB_CUDA(cudaStreamSynchronize(s));
  });
}

void mswb_aln_destroy(mswb_aln *aln) {
  if (!aln) return;
  cudaSetDevice(aln->ctx->device);
  delete aln;
}

} // extern "C"
