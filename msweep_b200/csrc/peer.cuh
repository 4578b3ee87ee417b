// peer.cuh — one-shot all-reduce over NVLink peer memory, done by ONE CTA inside our own kernels.
//
// per VI pass the ranks exchange K + 3 doubles (16 KB at K = 2000) — a latency-bound collective between
// multi-millisecond sweeps.  Calling NCCL for it costs a kernel launch of NCCL's plus a control kernel of ours behind
// it; here the CTA that owns the control step does the exchange itself:
//
//   push   every rank stores its vector into its own slot of every peer's receive area (plain remote stores over
//          NVLink / NVSwitch), fences at system scope, then releases one flag word per peer carrying the sequence number;
//   wait   it acquires the flag words its peers set in ITS area;
//   sum    it adds the world's vectors in RANK ORDER from local memory — every rank computes the same bits (no
//          dependence on a ring / tree order), so the redundant control steps stay bit-identical across ranks.
//
// Receive areas are double-buffered on the parity of the sequence number: a rank can be at most one collective ahead of
// a peer (it needs that peer's flag of collective s to finish s), so the buffers of parity s are never overwritten
// while a peer still reads them.  The sequence number lives on the device and advances only when a collective is
// EXECUTED, so kernels that exit early on every rank alike (converged, stalled) do not break the alternation.
// A peer that never arrives (crashed process, aborted context) ends the wait after a timeout or when the local abort
// word is set (mswb_ctx_abort); the caller then raises a fault instead of hanging.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mswb {

constexpr int PEER_MAX_WORLD = 16;
constexpr int PEER_SLOT_DOUBLES = 16384 + 64;     // one rank's vector: up to 16384 groups + the extra slots
constexpr size_t PEER_RECV_DOUBLES = (size_t)2 * PEER_MAX_WORLD * PEER_SLOT_DOUBLES;   // [parity][rank][slot]
constexpr int PEER_FLAG_WORDS = 2 * PEER_MAX_WORLD;                                     // [parity][rank]
// layout of one rank's block: recv area, then the flag words, then {seq, abort, error} (64-bit words)
constexpr size_t PEER_BLOCK_BYTES = PEER_RECV_DOUBLES * 8 + (size_t)PEER_FLAG_WORDS * 8 + 3 * 8;

struct PeerView {
  double *recv[PEER_MAX_WORLD];                 // recv[p]: rank p's receive area as mapped on THIS device
  unsigned long long *flags[PEER_MAX_WORLD];    // flags[p]: rank p's flag words
  unsigned long long *seq;                      // local: collectives executed so far
  unsigned long long *abort_word;               // local: set by mswb_ctx_abort
  unsigned long long *error_word;               // local: set when a wait gave up (read back by the host)
  unsigned long long timeout_ns;
  int world, rank;
};

#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long peer_ld_acquire(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void peer_st_release(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long peer_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// buf[0, count) := sum over ranks of their buf[0, count), in rank order.  All NT threads of ONE CTA per rank; every rank
// must call it the same number of times with the same count.  Returns false (in every thread) when a peer did not
// arrive — buf is then left as it was and *pv.error_word is set.
template <int NT>
__device__ __forceinline__ bool peer_allreduce_cta(double *buf, int count, const PeerView &pv) {
  __shared__ unsigned long long s_seq;
  __shared__ int s_bad;
  __syncthreads();
  if (threadIdx.x == 0) {
    s_seq = *pv.seq + 1ull;
    *pv.seq = s_seq;
    // once an exchange has given up, the ones enqueued behind it do not wait again
    s_bad = (*pv.error_word != 0ull || *(volatile unsigned long long *)pv.abort_word != 0ull) ? 1 : 0;
  }
  __syncthreads();
  if (s_bad) {
    if (threadIdx.x == 0 && *pv.error_word == 0ull) *pv.error_word = s_seq;
    return false;
  }
  const unsigned long long seq = s_seq;
  const int par = (int)(seq & 1ull);
  const size_t my_slot = ((size_t)par * PEER_MAX_WORLD + pv.rank) * PEER_SLOT_DOUBLES;
  // push (peers visited from rank + 1 on, so that the ranks do not all start on the same link)
  for (int d = 1; d < pv.world; ++d) {
    int p = pv.rank + d;
    if (p >= pv.world) p -= pv.world;
    double *dst = pv.recv[p] + my_slot;
    for (int i = threadIdx.x; i < count; i += NT) dst[i] = buf[i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < pv.world && (int)threadIdx.x != pv.rank)
    peer_st_release(pv.flags[threadIdx.x] + par * PEER_MAX_WORLD + pv.rank, seq);
  // wait for the peers' vectors
  if (threadIdx.x < pv.world && (int)threadIdx.x != pv.rank) {
    const unsigned long long *f = pv.flags[pv.rank] + par * PEER_MAX_WORLD + threadIdx.x;
    unsigned long long t0 = 0ull;
    unsigned spins = 0u;
    while (peer_ld_acquire(f) < seq) {
      if ((++spins & 255u) == 0u) {
        const unsigned long long now = peer_globaltimer();
        if (t0 == 0ull) t0 = now;
        if (now - t0 > pv.timeout_ns || *(volatile unsigned long long *)pv.abort_word != 0ull) { s_bad = 1; break; }
        __nanosleep(200);
      }
    }
  }
  __syncthreads();
  if (s_bad) {
    if (threadIdx.x == 0) *pv.error_word = seq;
    return false;
  }
  const double *mine = pv.recv[pv.rank] + (size_t)par * PEER_MAX_WORLD * PEER_SLOT_DOUBLES;
  for (int i = threadIdx.x; i < count; i += NT) {
    double s = 0.0;
    for (int q = 0; q < pv.world; ++q) s += q == pv.rank ? buf[i] : __ldcg(mine + (size_t)q * PEER_SLOT_DOUBLES + i);
    buf[i] = s;
  }
  __threadfence();
  __syncthreads();
  return true;
}
#endif

} // namespace mswb
