// handles.cuh — the opaque objects behind the C ABI (device-resident state).
#pragma once
#include "common.cuh"

// Equivalence-class table, replicated on every rank (include/mSWEEP_alignment.hpp:137-215 products).
struct mswb_aln {
  mswb_ctx *ctx = nullptr;
  uint64_t n_reads = 0, n_targets = 0, n_ecs = 0, n_aligned = 0, pat_nnz = 0;
  bool partitioned = false;           // true: the table covers only this rank's hash range of the reads
  mswb::DevBuf<uint64_t> hash;        // [n_ecs] ascending
  mswb::DevBuf<uint64_t> count;       // [n_ecs] reads per class
  mswb::DevBuf<uint32_t> rep_read;    // [n_ecs] smallest read id of the class
  mswb::DevBuf<uint64_t> pat_ptr;     // [n_ecs+1] CSR over the representative patterns
  mswb::DevBuf<uint32_t> pat_targets; // [pat_nnz]
  mswb::DevBuf<uint64_t> read_ptr;    // [n_ecs+1] CSR over member reads
  mswb::DevBuf<uint32_t> read_ids;    // [n_aligned] ascending inside each class
};

// Likelihood of the rank's EC shard plus the optimiser's persistent state.
//
// Device layout (private): EC-major.  Row j = class j of the shard, `Kp` elements per row
// (K rounded up so every row starts 16-byte aligned), padding columns hold 0.  This is the
// transpose of the reference's group-major seamat matrix: a VI pass then streams whole rows with
// 128-bit loads, the per-class logsumexp is a reduction along the row, and the per-group sums
// accumulate in registers down the rows.
struct mswb_lik {
  mswb_ctx *ctx = nullptr;
  uint32_t K_all = 0;          // groups in the grouping
  uint32_t K = 0;              // groups kept after --min-hits (rows of the reference's matrix)
  uint32_t Kp = 0;             // device row stride in elements
  uint64_t N = 0;              // classes in this rank's shard
  uint64_t N_pad = 0;          // N rounded up to ROW_PAD (common.cuh): counts / rowmax / P carry zero rows up to here
  uint64_t ec_begin = 0;       // first global class index of the shard
  uint64_t N_total = 0;        // classes over all ranks
  int storage = MSWB_STORE_F64;
  double sum_counts_total = 0; // sum of class counts over ALL ranks

  std::vector<uint8_t> mask;   // [K_all] groups_considered()
  std::vector<uint64_t> hits;  // [K_all] --min-hits tallies (empty when min_hits == 0)
  std::vector<uint32_t> kept;  // [K] original id of kept group g'

  // inputs kept for on-demand exports / rebuilds of the shard
  mswb::DevBuf<uint64_t> pat_ptr;          // [N+1] rebased to 0
  mswb::DevBuf<uint32_t> pat_targets;
  mswb::DevBuf<uint32_t> group_of_target;  // [T]
  mswb::DevBuf<uint32_t> kept_dev;         // [K]
  mswb::DevBuf<int> pos_dev;               // [K_all] row position of group g among the kept groups, -1 if pruned
  mswb::DevBuf<uint64_t> lut_off;          // [K+1] ragged LUT offsets
  mswb::DevBuf<double> lut;                // LUT[g'][c], c = 0..size(g')
  uint64_t n_targets = 0;
  bool from_patterns = false;
  double l0 = 0.0;                         // log(zero_inflation) = LUT[g][0] for every group

  mswb::DevBuf<double> counts;   // [N_pad] class counts c_j as doubles (0 in the padding)
  mswb::DevBuf<double> logl;     // [N x Kp] log-likelihood (RCG, exports); may be empty in F32 storage
  mswb::DevBuf<double> rowmax;   // [N] M_j = max_k logl(j, k)              (linear-domain EM)
  mswb::DevBuf<double> P64;      // [N x Kp] exp(logl - M_j)                 (EM, F64 storage)
  mswb::DevBuf<float> P32;       // [N x Kp4] same in fp32                   (EM, F32 storage)
  uint32_t Kp32 = 0;             // row stride of P32 (K rounded up to 4)
  // sparse storage: P(j,k) = P0[j] for the groups class j does not hit, P0[j] + dP for its hits (sorted by group)
  mswb::DevBuf<uint64_t> nz_ptr;   // [N+1]
  mswb::DevBuf<uint32_t> nz_grp;   // [nnz] row position (kept-group index) of the hit
  mswb::DevBuf<double> nz_dP;      // [nnz] exp(logl - M_j) - P0[j]
  mswb::DevBuf<double> P0;         // [N_pad] exp(log(zero_inflation) - M_j)
  mswb::DevBuf<double> nz_logl;    // [nnz] log-likelihood of the hit (sparse RCG works in the log domain)
  uint64_t nnz = 0;
  // sparse RCG state (vi_sparse_rcg.cuh): gamma = a_k + b_j off the hits, explicit on them; the same for the direction
  mswb::DevBuf<double> sp_b, sp_v; // [N] class part of gamma / of the search direction
  mswb::DevBuf<double> sp_g, sp_t; // [nnz] gamma / search direction of the hits

  // optimiser state that outlives a run (posterior export)
  mswb::DevBuf<double> gamma;    // [N x Kp] RCG log-responsibilities
  mswb::DevBuf<double> step;     // [N x Kp] RCG search direction
  mswb::DevBuf<double> last_dg;  // [K] digamma(N_k) of the last EM pass (posteriors on demand)
  mswb::DevBuf<double> last_a;   // [K] group part of gamma after the last sparse RCG run (posteriors on demand)
  int last_algo = -1;

  // result of the last mswb_vi_assign, waiting for mswb_vi_assign_fetch
  mswb::DevBuf<uint32_t> assign_reads;
  uint64_t assign_total = 0;
};

namespace mswb {
// (Re)build one device form of the likelihood matrix from the class patterns (likelihood.cu).
// On a matrix that takes a large part of HBM the other fp64 form is released first; it is rebuilt the
// same way when something asks for it again.
void lik_ensure_logl(mswb_lik *L);     // fp64 log-likelihood (RCG, exports)
void lik_ensure_linear(mswb_lik *L);   // P = exp(logl - rowmax) in the storage precision (EM)
void lik_ensure_sparse(mswb_lik *L);   // hit lists (MSWB_STORE_SPARSE)
// K x n tile of log-posteriors of the last run for classes [ec_begin, ec_begin + n), on the device (vi.cu)
void posterior_tile_dev(mswb_ctx *ctx, mswb_lik *lik, uint64_t ec_begin, uint64_t n, double *tile);
} // namespace mswb
