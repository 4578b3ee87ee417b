"""msweep_b200 — B200-native abundance-estimation backend for mSWEEP.

The product is the C-ABI CUDA library `lib/libmsweep_b200.so` (declared in include/msweep_b200.h) and
the C++17 host driver `bin/mSWEEP_b200`.  This module is only the ctypes view of that C ABI used by
the tests and bench.py: every call below goes straight through the exported `mswb_*` symbols.
There is no CPU fallback and no Python implementation of anything on the path: if the library is
missing or no GPU is present, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MSWB_LIB_PATH", os.path.join(_HERE, "lib", "libmsweep_b200.so"))   # override: A/B runs only
HEADER_PATH = os.path.join(_HERE, "..", "include", "msweep_b200.h")

ALGO_RCG, ALGO_EM = 0, 1
STORE_F64, STORE_F32, STORE_SPARSE = 0, 1, 2
RNG_EXACT, RNG_PHILOX = 0, 1
NCCL_ID_BYTES = 128


class MswbError(RuntimeError):
    pass


class ViOpts(C.Structure):
    _fields_ = [("tol", C.c_double), ("max_iters", C.c_uint64), ("algo", C.c_int), ("time_kernels", C.c_int),
                ("poll_every", C.c_uint32)]


class ViStat(C.Structure):
    _fields_ = [("bound", C.c_double), ("gnorm", C.c_double), ("iters", C.c_uint64), ("converged", C.c_int),
                ("resets", C.c_uint64), ("pass_ms_sum", C.c_double), ("pass_launches", C.c_uint64),
                ("pass_bytes", C.c_uint64)]


_lib = None


def lib() -> C.CDLL:
    """Loads libmsweep_b200.so; raises if it has not been built (python -m msweep_b200._build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MswbError(f"{LIB_PATH} is missing: build it with `python -m msweep_b200._build` "
                            "(there is no fallback implementation)")
        L = C.CDLL(LIB_PATH)
        L.mswb_last_error.restype = C.c_char_p
        L.mswb_version.restype = C.c_char_p
        L.mswb_launch_count.restype = C.c_uint64
        L.mswb_pattern_hash.restype = C.c_uint64
        L.mswb_pattern_hash.argtypes = [C.c_void_p, C.c_uint64]
        _lib = L
    return _lib


def _check(rc: int) -> None:
    if rc != 0:
        raise MswbError(lib().mswb_last_error().decode())


def _p(arr, ctype=C.c_void_p):
    return None if arr is None else arr.ctypes.data_as(ctype)


def pattern_hash(targets) -> int:
    """mswb_pattern_hash of one ascending target list (host-side; for hash-range partitioning of reads)."""
    t = np.ascontiguousarray(targets, np.uint32)
    return int(lib().mswb_pattern_hash(t.ctypes.data if len(t) else None, len(t)))


def mt64_jump(state, n_outputs: int) -> np.ndarray:
    """mswb_mt64_jump: the 312 words of std::mt19937_64 (at a refill boundary) n_outputs outputs later (host arithmetic)."""
    st = np.ascontiguousarray(state, np.uint64)
    assert st.shape == (312,)
    out = np.zeros(312, np.uint64)
    _check(lib().mswb_mt64_jump(st.ctypes.data_as(C.c_void_p), C.c_uint64(n_outputs), out.ctypes.data_as(C.c_void_p)))
    return out


def launch_count() -> int:
    return int(lib().mswb_launch_count())


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(NCCL_ID_BYTES)
    _check(lib().mswb_nccl_unique_id(buf))
    return buf.raw


class Context:
    def __init__(self, device: int = 0, rank: int = 0, world_size: int = 1, nccl_id: bytes | None = None,
                 cuda_stream: int | None = None):
        self.h = C.c_void_p()
        self.rank, self.world_size, self.device = rank, world_size, device
        _check(lib().mswb_ctx_create(device, rank, world_size, nccl_id, C.c_void_p(cuda_stream or 0), C.byref(self.h)))

    def sync(self) -> None:
        _check(lib().mswb_ctx_sync(self.h))

    def trim(self) -> None:
        """mswb_ctx_trim: hand the parked device blocks of this context's GPU back to the driver."""
        _check(lib().mswb_ctx_trim(self.h))

    @property
    def peer_active(self) -> bool:
        """mswb_ctx_peer_active: the per-pass all-reduce runs over NVLink peer memory inside our own kernel (else NCCL)."""
        return bool(lib().mswb_ctx_peer_active(self.h))

    def abort(self) -> None:
        """mswb_ctx_abort: abort the NCCL communicator (a peer failed); pending collectives end with an error."""
        _check(lib().mswb_ctx_abort(self.h))

    def shard_range(self, n_ecs: int) -> tuple[int, int]:
        a, b = C.c_uint64(), C.c_uint64()
        _check(lib().mswb_shard_range(self.h, C.c_uint64(n_ecs), C.byref(a), C.byref(b)))
        return a.value, b.value

    def close(self) -> None:
        if self.h:
            lib().mswb_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


@dataclass
class EcExport:
    hash: np.ndarray
    count: np.ndarray
    rep_read: np.ndarray
    pat_ptr: np.ndarray
    pat_targets: np.ndarray
    read_ptr: np.ndarray
    read_ids: np.ndarray


class Alignment:
    """mswb_ec_build: equivalence classes from the strand-merged pseudoalignment (CSR over reads)."""

    def __init__(self, ctx: Context, n_reads: int, n_targets: int, row_ptr: np.ndarray, targets: np.ndarray,
                 partitioned: bool = False):
        self.ctx = ctx
        self.h = C.c_void_p()
        rp = np.ascontiguousarray(row_ptr, np.uint64)
        tg = np.ascontiguousarray(targets, np.uint32)
        assert len(rp) == n_reads + 1
        fn = lib().mswb_ec_build_partitioned if partitioned else lib().mswb_ec_build
        _check(fn(ctx.h, C.c_uint64(n_reads), C.c_uint64(n_targets), _p(rp), _p(tg) if len(tg) else None,
                                   C.byref(self.h)))
        v = [C.c_uint64() for _ in range(4)]
        _check(lib().mswb_ec_info(self.h, *[C.byref(x) for x in v]))
        self.n_ecs, self.n_reads, self.n_aligned, self.pat_nnz = [x.value for x in v]
        self.n_targets = n_targets

    def export(self) -> EcExport:
        n = self.n_ecs
        e = EcExport(np.zeros(n, np.uint64), np.zeros(n, np.uint64), np.zeros(n, np.uint32), np.zeros(n + 1, np.uint64),
                     np.zeros(self.pat_nnz, np.uint32), np.zeros(n + 1, np.uint64), np.zeros(self.n_aligned, np.uint32))
        _check(lib().mswb_ec_export(self.h, _p(e.hash), _p(e.count), _p(e.rep_read), _p(e.pat_ptr), _p(e.pat_targets),
                                    _p(e.read_ptr), _p(e.read_ids)))
        return e

    def close(self) -> None:
        if self.h:
            lib().mswb_aln_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


@dataclass
class ViResult:
    theta: np.ndarray
    N_k: np.ndarray | None
    bound: float
    gnorm: float
    iters: int
    converged: bool
    resets: int
    pass_ms_sum: float = 0.0
    pass_launches: int = 0
    pass_bytes: int = 0
    trace_bound: np.ndarray | None = None
    trace_gnorm: np.ndarray | None = None
    trace_reset: np.ndarray | None = None


class Likelihood:
    """mswb_lik_build / mswb_lik_from_dense + the optimiser entry points that run on it."""

    def __init__(self, ctx: Context, handle: C.c_void_p):
        self.ctx, self.h = ctx, handle
        a, b = C.c_uint32(), C.c_uint32()
        c, d, e = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _check(lib().mswb_lik_info(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d), C.byref(e)))
        self.n_groups_all, self.n_groups, self.n_ecs, self.ec_begin, self.n_ecs_total = a.value, b.value, c.value, d.value, e.value

    @classmethod
    def build(cls, ctx: Context, aln: Alignment, group_of_target, group_sizes, q=0.65, e=0.01, zero_inflation=0.01,
              min_hits=0, storage=STORE_F64) -> "Likelihood":
        got = np.ascontiguousarray(group_of_target, np.uint32)
        sz = np.ascontiguousarray(group_sizes, np.uint64)
        assert len(got) == aln.n_targets
        h = C.c_void_p()
        _check(lib().mswb_lik_build(ctx.h, aln.h, _p(got), C.c_uint32(len(sz)), _p(sz), C.c_double(q), C.c_double(e),
                                    C.c_double(zero_inflation), C.c_uint64(min_hits), C.c_int(storage), C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def from_dense(cls, ctx: Context, logl: np.ndarray, log_counts: np.ndarray, storage=STORE_F64) -> "Likelihood":
        """logl: (K, N_local) group-major, as the reference's seamat matrix."""
        m = np.ascontiguousarray(logl, np.float64)
        lc = np.ascontiguousarray(log_counts, np.float64)
        K, N = m.shape
        assert len(lc) == N
        h = C.c_void_p()
        _check(lib().mswb_lik_from_dense(ctx.h, _p(m), C.c_uint32(K), C.c_uint64(N), _p(lc), C.c_int(storage), C.byref(h)))
        return cls(ctx, h)

    def mask(self, want_hits: bool = False):
        m = np.zeros(self.n_groups_all, np.uint8)
        hits = np.zeros(self.n_groups_all, np.uint64) if want_hits else None
        _check(lib().mswb_lik_mask(self.h, _p(m), _p(hits)))
        return (m, hits) if want_hits else m

    def export_hit_counts(self) -> np.ndarray:
        out = np.zeros((self.n_groups_all, self.n_ecs), np.uint32)
        _check(lib().mswb_lik_export_hit_counts(self.h, _p(out)))
        return out

    def export_logl(self) -> np.ndarray:
        out = np.zeros((self.n_groups, self.n_ecs), np.float64)
        _check(lib().mswb_lik_export_logl(self.h, _p(out)))
        return out

    # ---- optimiser -----------------------------------------------------------------------------
    def vi_run(self, algo=ALGO_RCG, alpha0=None, log_counts=None, tol=1e-6, max_iters=5000, time_kernels=False,
               poll_every=0, on_iter=None) -> ViResult:
        K = self.n_groups
        a0 = np.ones(K) if alpha0 is None else np.ascontiguousarray(alpha0, np.float64)
        assert len(a0) == K
        lc = None if log_counts is None else np.ascontiguousarray(log_counts, np.float64)
        opts = ViOpts(tol, max_iters, algo, int(time_kernels), poll_every)
        stat = ViStat()
        theta = np.zeros(K)
        CB = C.CFUNCTYPE(None, C.c_void_p, C.c_uint64, C.c_double, C.c_double)
        cb = CB(lambda user, it, b, g: on_iter(it, b, g)) if on_iter else C.cast(None, CB)
        _check(lib().mswb_vi_run(self.ctx.h, self.h, _p(a0), _p(lc), C.byref(opts), _p(theta), C.byref(stat), cb, None))
        return _result(theta, None, stat)

    def vi_begin(self, algo=ALGO_RCG, alpha0=None, log_counts=None, tol=1e-6, max_iters=5000, time_kernels=False) -> "ViSession":
        return ViSession(self, algo, alpha0, log_counts, tol, max_iters, time_kernels)

    def posteriors(self, ec_begin: int = 0, ec_end: int | None = None) -> np.ndarray:
        ec_end = self.n_ecs if ec_end is None else ec_end
        out = np.zeros((self.n_groups, ec_end - ec_begin))
        _check(lib().mswb_vi_posteriors(self.ctx.h, self.h, C.c_uint64(ec_begin), C.c_uint64(ec_end), _p(out)))
        return out

    def assign(self, aln: "Alignment", log_threshold) -> list[np.ndarray]:
        """mswb_vi_assign + _fetch: bins[k] = ascending read ids of the local classes with log-posterior >= log_threshold[k]."""
        thr = np.ascontiguousarray(log_threshold, np.float64)
        assert thr.shape == (self.n_groups,)
        ptr = np.zeros(self.n_groups + 1, np.uint64)
        _check(lib().mswb_vi_assign(self.ctx.h, self.h, aln.h, _p(thr), _p(ptr)))
        flat = np.zeros(max(1, int(ptr[-1])), np.uint32)
        _check(lib().mswb_vi_assign_fetch(self.h, _p(flat)))
        return [flat[int(ptr[k]):int(ptr[k + 1])].copy() for k in range(self.n_groups)]

    # ---- bootstrap -----------------------------------------------------------------------------
    def bootstrap_resample(self, seed: int, n_replicates: int, bootstrap_count: int = 0, rng_mode=RNG_EXACT) -> np.ndarray:
        out = np.zeros((n_replicates, self.n_ecs_total), np.uint32)
        _check(lib().mswb_bootstrap_resample(self.ctx.h, self.h, C.c_int32(seed), C.c_uint64(bootstrap_count),
                                             C.c_int(rng_mode), C.c_uint64(n_replicates), _p(out)))
        return out

    def bootstrap_run(self, n_replicates: int, seed: int, algo=ALGO_RCG, alpha0=None, tol=1e-6, max_iters=5000,
                      bootstrap_count: int = 0, rng_mode=RNG_EXACT, replica_rank: int = 0, replica_world: int = 1):
        K = self.n_groups
        a0 = np.ones(K) if alpha0 is None else np.ascontiguousarray(alpha0, np.float64)
        opts = ViOpts(tol, max_iters, algo, 0, 0)
        thetas = np.full((n_replicates, K), np.nan)
        stats = (ViStat * n_replicates)()
        _check(lib().mswb_bootstrap_run(self.ctx.h, self.h, _p(a0), C.byref(opts), C.c_uint64(n_replicates),
                                        C.c_uint64(bootstrap_count), C.c_int32(seed), C.c_int(rng_mode),
                                        C.c_int(replica_rank), C.c_int(replica_world), _p(thetas), stats))
        return thetas, [int(s.iters) for s in stats]

    def close(self) -> None:
        if self.h:
            lib().mswb_lik_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _result(theta, N_k, stat: ViStat) -> ViResult:
    return ViResult(theta, N_k, stat.bound, stat.gnorm, int(stat.iters), bool(stat.converged), int(stat.resets),
                    stat.pass_ms_sum, int(stat.pass_launches), int(stat.pass_bytes))


class ViSession:
    """mswb_vi_begin / _step / _poll / _finish: the stepwise interface bench.py times."""

    def __init__(self, lik: Likelihood, algo, alpha0, log_counts, tol, max_iters, time_kernels):
        self.lik = lik
        K = lik.n_groups
        a0 = np.ones(K) if alpha0 is None else np.ascontiguousarray(alpha0, np.float64)
        lc = None if log_counts is None else np.ascontiguousarray(log_counts, np.float64)
        self.opts = ViOpts(tol, max_iters, algo, int(time_kernels), 0)
        self.h = C.c_void_p()
        self.max_iters = max_iters
        _check(lib().mswb_vi_begin(lik.ctx.h, lik.h, _p(a0), _p(lc), C.byref(self.opts), C.byref(self.h)))

    def step(self, n_iters: int) -> None:
        _check(lib().mswb_vi_step(self.h, C.c_uint64(n_iters)))

    def poll(self) -> ViStat:
        stat = ViStat()
        _check(lib().mswb_vi_poll(self.h, C.byref(stat)))
        return stat

    def trace(self):
        cap = min(self.max_iters, 1 << 20)
        b, g, r = np.zeros(cap), np.zeros(cap), np.zeros(cap, np.uint8)
        _check(lib().mswb_vi_trace(self.h, _p(b), _p(g), _p(r), C.c_uint64(cap)))
        n = int(self.poll().iters)
        return b[:n], g[:n], r[:n]

    def finish(self) -> ViResult:
        K = self.lik.n_groups
        theta, Nk = np.zeros(K), np.zeros(K)
        stat = ViStat()
        h, self.h = self.h, C.c_void_p()
        _check(lib().mswb_vi_finish(h, _p(theta), _p(Nk), C.byref(stat)))
        return _result(theta, Nk, stat)


def declared_symbols() -> list[str]:
    """Every function name include/msweep_b200.h declares (used by the CPU-side ABI test)."""
    import re
    with open(HEADER_PATH) as f:
        txt = f.read()
    return sorted(set(re.findall(r"MSWB_API[^;(]*?\b(mswb_[a-z0-9_]+)\s*\(", txt)))
