/* msweep_b200.h — C ABI of the B200-native abundance-estimation backend for mSWEEP.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  It replaces the
 * LibTorch `rcggpu`/`emgpu` path of the reference, i.e. exactly the work the reference does between
 * `alignment.collapse()` and `write_abundances()`:
 *
 *   reference (file:line, relative to the mSWEEP tree)                     entry point here
 *   -------------------------------------------------------------------    -------------------------
 *   Alignment::collapse            include/mSWEEP_alignment.hpp:137-215     mswb_ec_build
 *   LL_WOR21::fill_ll_mat          include/Likelihood.hpp:109-186           mswb_lik_build
 *   LL_WOR21::fill_ec_counts       include/Likelihood.hpp:188-195           mswb_lik_build
 *   Likelihood::groups_considered  include/Likelihood.hpp:79, 330           mswb_lik_mask
 *   Likelihood::from_file / log_mat include/Likelihood.hpp:224-252, 321-324 mswb_lik_from_dense
 *   rcg_optl -> rcgpar::rcg_optl_omp / rcg_optl_torch  src/mSWEEP.cpp:192-199   mswb_vi_run (MSWB_ALGO_RCG)
 *   rcg_optl -> rcgpar::em_torch   src/mSWEEP.cpp:200-203                   mswb_vi_run (MSWB_ALGO_EM)
 *   rcgpar::mixture_components[_torch]  src/mSWEEP.cpp:419-423, 512-516     theta output of mswb_vi_run
 *   Sample::store_probs / write_probs   src/Sample.cpp:63-85                mswb_vi_posteriors (on demand)
 *   mGEMS::BinFromMatrix hand-off       src/mSWEEP.cpp:437-469              mswb_vi_assign / _fetch
 *   BootstrapSample::resample_counts    src/BootstrapSample.cpp:60-73       mswb_bootstrap_resample
 *   bootstrap loop                 src/mSWEEP.cpp:496-518                   mswb_bootstrap_run
 *
 * Conventions
 *   - Every function returns 0 on success, non-zero on failure; mswb_last_error() then gives the
 *     message (thread-local).  A host shim turns that into the std::exception the reference's
 *     catch sites expect (src/mSWEEP.cpp:400-406, 506-511).
 *   - All pointer arguments are HOST pointers unless the name ends in `_dev`.
 *   - One mswb_ctx drives ONE GPU.  Multi-GPU runs use one ctx per GPU (one process per GPU under
 *     torchrun, or one host thread per GPU), joined by an NCCL communicator: equivalence classes are
 *     sharded by contiguous ranges and every VI pass all-reduces K+2 doubles.
 *   - Matrices cross this boundary in the reference's orientation: K groups x N classes, group-major
 *     (row = group), exactly what seamat::DenseMatrix<double> holds.  The device layout is private.
 *   - There is no CPU fallback: without a CUDA device every compute entry fails with an error.
 */
#ifndef MSWEEP_B200_H
#define MSWEEP_B200_H

#include <stdint.h>

#if defined(__GNUC__)
#define MSWB_API __attribute__((visibility("default")))
#else
#define MSWB_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mswb_ctx mswb_ctx;   /* device, stream, NCCL communicator, scratch            */
typedef struct mswb_aln mswb_aln;   /* device-resident equivalence-class table + patterns    */
typedef struct mswb_lik mswb_lik;   /* device-resident (EC-sharded) likelihood + VI state    */
typedef struct mswb_vi  mswb_vi;    /* one running optimisation (stepwise interface)         */

MSWB_API const char *mswb_last_error(void);
MSWB_API const char *mswb_version(void);
/* Number of kernels of this library launched by the calling process so far. */
MSWB_API uint64_t mswb_launch_count(void);

/* ---- context ------------------------------------------------------------------------------- */
#define MSWB_NCCL_ID_BYTES 128
/* Rank 0 creates the id and ships it to the other ranks (any transport). */
MSWB_API int  mswb_nccl_unique_id(void *out_id /* MSWB_NCCL_ID_BYTES */);
/* cuda_stream: a cudaStream_t to enqueue on (so a caller can bracket work with its own events),
 * or NULL for a private non-blocking stream.  nccl_id may be NULL iff world_size == 1. */
MSWB_API int  mswb_ctx_create(int device, int rank, int world_size, const void *nccl_id, void *cuda_stream,
                     mswb_ctx **out);
/* Single-process form: one context per listed device, rank i on devices[i], joined by one NCCL clique
 * (ncclCommInitAll — no bootstrap network, so it is much quicker than n calls of mswb_ctx_create).  out[n]. */
MSWB_API int  mswb_ctx_create_group(int n, const int *devices, mswb_ctx **out);
MSWB_API void mswb_ctx_destroy(mswb_ctx *ctx);
/* Device blocks of 64 MB and more that the library has released are parked for reuse (cudaMalloc / cudaFree of a
 * 100 GB matrix cost 0.2-0.4 s), up to MSWB_CACHE_GB gigabytes per device (default: a quarter of its memory).
 * An embedding process that needs the memory back calls this: every parked block of ctx's device is cudaFree'd. */
MSWB_API int  mswb_ctx_trim(mswb_ctx *ctx);
/* Aborts the NCCL communicator of ctx (ncclCommAbort): collectives this rank is waiting in end with an error instead
 * of waiting for a peer that has failed.  May be called from another host thread than the one driving ctx; a process
 * that holds several ranks calls it on every context when one of them fails.  The context can still be destroyed. */
MSWB_API int  mswb_ctx_abort(mswb_ctx *ctx);
/* 1 when the per-pass all-reduce of this context (K + 3 doubles per EM pass, a scalar and K + 1 per RCG
 * iteration) runs as a one-shot exchange over NVLink peer memory inside the library's own control kernel — every rank
 * maps every other rank's receive block at context creation (CUDA IPC between processes, peer access between the
 * contexts of one process) and sums the world's vectors in rank order; 0 when NCCL carries it (one GPU, no peer access
 * between some pair of devices, or MSWB_PEER=0 in the environment).  The decision is taken collectively: all ranks of
 * a world answer the same.  mswb_ctx_abort also ends a wait on a peer that will never arrive. */
MSWB_API int  mswb_ctx_peer_active(const mswb_ctx *ctx);
/* Creates the CUDA primary context of `device` (seconds on a cold process); call it from a side thread while
 * the host is still parsing its inputs so that mswb_ctx_create returns immediately afterwards. */
MSWB_API int  mswb_device_warmup(int device);
MSWB_API int  mswb_ctx_sync(mswb_ctx *ctx);
/* Contiguous EC range [*begin, *end) owned by this rank when n_ecs classes are sharded. */
MSWB_API int  mswb_shard_range(const mswb_ctx *ctx, uint64_t n_ecs, uint64_t *begin, uint64_t *end);

/* ---- (1a) equivalence classes -------------------------------------------------------------- */
/* Input: the strand-merged pseudoalignment as CSR, row r = ascending target ids of read r
 * (an empty row = unaligned read, counted in n_reads but in no class).  Replicated on every rank. */
MSWB_API int  mswb_ec_build(mswb_ctx *ctx, uint64_t n_reads, uint64_t n_targets, const uint64_t *row_ptr,
                   const uint32_t *targets, mswb_aln **out);
/* Same, but the reads given are ONLY this rank's partition of the alignment: the host has routed reads
 * to ranks by ascending ranges of the pattern hash (mswb_pattern_hash), rank 0 owning the lowest range.
 * The global class table is then the concatenation of the ranks' tables, no class straddles two
 * ranks, and mswb_lik_build keeps the whole local table as this rank's shard. */
MSWB_API int  mswb_ec_build_partitioned(mswb_ctx *ctx, uint64_t n_reads_local, uint64_t n_targets,
                               const uint64_t *row_ptr, const uint32_t *targets, mswb_aln **out);
/* The reference's pattern hash (include/mSWEEP_alignment.hpp:150-155) of one ascending target list,
 * for hosts that partition reads by hash range. */
MSWB_API uint64_t mswb_pattern_hash(const uint32_t *targets, uint64_t n);
MSWB_API int  mswb_ec_info(const mswb_aln *aln, uint64_t *n_ecs, uint64_t *n_reads, uint64_t *n_aligned,
                  uint64_t *pattern_nnz);
/* Parity export; any pointer may be NULL.  hash/count/rep_read: n_ecs; pat_ptr/read_ptr: n_ecs+1;
 * pat_targets: pattern_nnz; read_ids: n_aligned (ascending inside each class). */
MSWB_API int  mswb_ec_export(const mswb_aln *aln, uint64_t *hash, uint64_t *count, uint32_t *rep_read,
                    uint64_t *pat_ptr, uint32_t *pat_targets, uint64_t *read_ptr, uint32_t *read_ids);
MSWB_API void mswb_aln_destroy(mswb_aln *aln);

/* ---- (1b) likelihood ----------------------------------------------------------------------- */
/* F64 / F32: the dense K x N matrix the reference builds (fp32 = --emprecision float, EM only).
 * SPARSE (EM only, from class patterns only): LL_WOR21 gives every group a class does NOT hit the same value
 * log(zero_inflation), so a class is stored as that constant plus its few (group, value) hits in fp64 — the same
 * numbers as F64, O(hits) instead of O(K) bytes per class; config 3 (1e8 x 2000) then fits one GPU. */
enum { MSWB_STORE_F64 = 0, MSWB_STORE_F32 = 1, MSWB_STORE_SPARSE = 2 };
/* group_of_target[n_targets], group_sizes[n_groups] as read from the -i file (ids in order of first
 * appearance).  q, e, zero_inflation, min_hits are the reference's -q, -e, --zero-inflation,
 * --min-hits.  The rank keeps the rows of its own EC shard. */
MSWB_API int  mswb_lik_build(mswb_ctx *ctx, const mswb_aln *aln, const uint32_t *group_of_target,
                    uint32_t n_groups, const uint64_t *group_sizes, double q, double e,
                    double zero_inflation, uint64_t min_hits, int storage, mswb_lik **out);
/* A precomputed likelihood (parity tests, --read-likelihood): logl is K x N_local group-major,
 * log_counts[N_local] natural logs of the class counts; each rank passes its own shard. */
MSWB_API int  mswb_lik_from_dense(mswb_ctx *ctx, const double *logl, uint32_t n_groups, uint64_t n_ecs_local,
                         const double *log_counts, int storage, mswb_lik **out);
MSWB_API int  mswb_lik_info(const mswb_lik *lik, uint32_t *n_groups_all, uint32_t *n_groups_kept,
                   uint64_t *n_ecs_local, uint64_t *ec_begin, uint64_t *n_ecs_total);
/* groups_considered() and the --min-hits tallies; either pointer may be NULL. */
MSWB_API int  mswb_lik_mask(const mswb_lik *lik, uint8_t *mask /* n_groups_all */, uint64_t *hits /* n_groups_all */);
/* Parity exports of the local shard: hit counts c(g,i) (n_groups_all x n_ecs_local, group-major)
 * and the log-likelihood matrix (n_groups_kept x n_ecs_local, group-major). */
MSWB_API int  mswb_lik_export_hit_counts(const mswb_lik *lik, uint32_t *out);
MSWB_API int  mswb_lik_export_logl(const mswb_lik *lik, double *out);
MSWB_API void mswb_lik_destroy(mswb_lik *lik);

/* ---- (2) variational inference ------------------------------------------------------------- */
enum { MSWB_ALGO_RCG = 0, MSWB_ALGO_EM = 1 };
typedef struct {
  double   tol;          /* --tol       (reference default 1e-6, src/mSWEEP.cpp:125)               */
  uint64_t max_iters;    /* --max-iters (reference default 5000, src/mSWEEP.cpp:123)               */
  int      algo;         /* MSWB_ALGO_RCG | MSWB_ALGO_EM                                           */
  int      time_kernels; /* 1: record CUDA events around every pass kernel (see mswb_vi_stat)      */
  uint32_t poll_every;   /* iterations enqueued between host convergence polls (0 = default)      */
} mswb_vi_opts;
typedef struct {
  double   bound;        /* ELBO at exit                                                           */
  double   gnorm;        /* last Riemannian gradient norm (RCG), 0 for EM                          */
  uint64_t iters;        /* iterations executed                                                    */
  int      converged;
  uint64_t resets;       /* RCG restarts taken                                                     */
  double   pass_ms_sum;  /* time_kernels: summed device time of the pass kernels, and their count  */
  uint64_t pass_launches;
  uint64_t pass_bytes;   /* algorithmic HBM bytes of the sweep(s) of ONE iteration on this rank      */
} mswb_vi_stat;
typedef void (*mswb_iter_cb)(void *user, uint64_t iter, double bound, double gnorm);

/* alpha0[K_kept] Dirichlet pseudo-counts (src/mSWEEP.cpp:391-398).  log_counts: NULL to use the class
 * counts held by lik, else n_ecs_local natural logs (-inf allowed = class not observed; bootstrap).
 * theta[K_kept] receives the relative abundances (identical on every rank). */
MSWB_API int  mswb_vi_run(mswb_ctx *ctx, mswb_lik *lik, const double *alpha0, const double *log_counts,
                 const mswb_vi_opts *opts, double *theta, mswb_vi_stat *stat, mswb_iter_cb on_iter, void *user);

/* Stepwise form of the same thing (what mswb_vi_run is built from; bench.py times _step). */
MSWB_API int  mswb_vi_begin(mswb_ctx *ctx, mswb_lik *lik, const double *alpha0, const double *log_counts,
                   const mswb_vi_opts *opts, mswb_vi **out);
MSWB_API int  mswb_vi_step(mswb_vi *vi, uint64_t n_iters);               /* enqueue only, no host sync   */
MSWB_API int  mswb_vi_poll(mswb_vi *vi, mswb_vi_stat *stat);             /* sync + read device status    */
MSWB_API int  mswb_vi_trace(mswb_vi *vi, double *bound, double *gnorm, uint8_t *reset, uint64_t capacity);
MSWB_API int  mswb_vi_finish(mswb_vi *vi, double *theta, double *N_k, mswb_vi_stat *stat);   /* frees vi */

/* log-posteriors of the LAST run for local classes [ec_begin, ec_end) (local indices), written as
 * K_kept x (ec_end-ec_begin) group-major — what rcg_optl returns, tile by tile. */
MSWB_API int  mswb_vi_posteriors(mswb_ctx *ctx, mswb_lik *lik, uint64_t ec_begin, uint64_t ec_end, double *gamma);

/* Thresholded class -> group assignment, the hand-off to read binning (src/mSWEEP.cpp:437-469 ->
 * mGEMS::BinFromMatrix with Alignment::get_aligned_reads, include/mSWEEP_alignment.hpp:241; mGEMS v1.3.3 is
 * off-tree, rule restated from its published description).  Local class j — all of its reads — joins the
 * bin of kept group k when the LAST run's log-posterior gamma(k, j) >= log_threshold[k] (mGEMS: the log of
 * the group's abundance; pass +inf to skip a group).  bin_ptr (K_kept + 1) receives the CSR offsets; the
 * member read ids (ascending inside each bin) wait on the device for mswb_vi_assign_fetch, which copies
 * bin_ptr[K_kept] ids out and releases them.  `aln` is the alignment the likelihood was built from. */
MSWB_API int  mswb_vi_assign(mswb_ctx *ctx, mswb_lik *lik, const mswb_aln *aln, const double *log_threshold,
                    uint64_t *bin_ptr);
MSWB_API int  mswb_vi_assign_fetch(mswb_lik *lik, uint32_t *read_ids);

/* ---- (3) bootstrap ------------------------------------------------------------------------- */
enum { MSWB_RNG_LIBSTDCXX_EXACT = 0, MSWB_RNG_PHILOX = 1 };
/* Resampled class counts of n_replicates consecutive replicates (out: n_replicates x n_ecs_total),
 * the integer vector BootstrapSample::resample_counts takes the log of. */
MSWB_API int  mswb_bootstrap_resample(mswb_ctx *ctx, const mswb_lik *lik, int32_t seed, uint64_t bootstrap_count,
                             int rng_mode, uint64_t n_replicates, uint32_t *out);
/* n_replicates full re-estimations; replicate r runs on the rank r % world_size (each rank must
 * hold the whole likelihood: build it with world_size == 1 contexts, or see DESIGN.md).
 * thetas: n_replicates x K_kept, rows of other ranks are left untouched. */
MSWB_API int  mswb_bootstrap_run(mswb_ctx *ctx, mswb_lik *lik, const double *alpha0, const mswb_vi_opts *opts,
                        uint64_t n_replicates, uint64_t bootstrap_count, int32_t seed, int rng_mode,
                        int replica_rank, int replica_world, double *thetas, mswb_vi_stat *stats);

/* Jump-ahead of std::mt19937_64 (host arithmetic, no device work).  The reference seeds ONE generator and draws all
 * replicates from it (src/BootstrapSample.cpp:48-60), so replicate r starts r * bootstrap_count outputs into the stream;
 * mswb_bootstrap_run uses this to stand a rank at the start of its own replicates without producing the others' draws.
 * state[312]: the generator's words with all 312 consumed (as right after seeding); state_out[312]: the same n_outputs
 * outputs later.  The low 31 bits of state_out[0] are not part of the generator's state. */
MSWB_API int  mswb_mt64_jump(const uint64_t *state, uint64_t n_outputs, uint64_t *state_out);

#ifdef __cplusplus
}
#endif
#endif /* MSWEEP_B200_H */
