"""ctypes binding for the CPU oracle (oracle/liboracle.so) and, when built, for the piece of the
real reference that compiles here (oracle/_ref/libmsweep_ref_grouping.so).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs; never by the product package msweep_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def build(force: bool = False) -> None:
    """Compile the oracle with its Makefile (and oracle/_ref when /root/reference is present)."""
    if force or not os.path.exists(os.path.join(_HERE, "liboracle.so")):
        subprocess.check_call(["make", "-C", _HERE, "-j4"], stdout=subprocess.DEVNULL)


def _opt(arr):
    return None if arr is None else arr.ctypes.data_as(C.c_void_p)


class _Lib:
    def __init__(self, name: str):
        build()
        self.lib = C.CDLL(os.path.join(_HERE, name))
        L = self.lib
        L.orc_last_error.restype = C.c_char_p
        L.orc_pattern_hash.restype = C.c_uint64
        L.orc_pattern_hash.argtypes = [_u32p, C.c_uint64]
        L.orc_digamma.restype = C.c_double
        L.orc_digamma.argtypes = [C.c_double]
        L.orc_ldbb_scaled.restype = C.c_double
        L.orc_ldbb_scaled.argtypes = [C.c_uint64, C.c_uint64, C.c_double, C.c_double]
        L.orc_bb_parameters.argtypes = [C.c_uint64, C.c_double, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        for f in ("orc_reads_n", "orc_reads_nnz", "orc_ec_n", "orc_ec_pat_nnz", "orc_ec_n_read_ids",
                  "orc_grouping_n_targets", "orc_lik_n_ecs", "orc_lik_lut_cols"):
            getattr(L, f).restype = C.c_uint64
            getattr(L, f).argtypes = [C.c_void_p]
        for f in ("orc_grouping_n_groups", "orc_lik_n_groups"):
            getattr(L, f).restype = C.c_uint32
            getattr(L, f).argtypes = [C.c_void_p]
        for f in ("orc_reads_free", "orc_ec_free", "orc_grouping_free", "orc_lik_free"):
            getattr(L, f).argtypes = [C.c_void_p]
            getattr(L, f).restype = None
        L.orc_grouping_name.restype = C.c_char_p
        L.orc_grouping_name.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_reads_from_csr.argtypes = [C.c_uint64, C.c_uint64, _u64p, _u32p, C.POINTER(C.c_void_p)]
        L.orc_reads_csr.argtypes = [C.c_void_p, _u64p, _u32p]
        L.orc_ec_build.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        L.orc_ec_export.argtypes = [C.c_void_p] + [C.c_void_p] * 7
        L.orc_grouping_export.argtypes = [C.c_void_p, _u32p, _u64p]
        L.orc_lik_build.argtypes = [C.c_void_p, _u32p, C.c_uint64, C.c_uint32, _u64p, C.c_double, C.c_double,
                                    C.c_double, C.c_uint64, C.c_int, C.POINTER(C.c_void_p)]
        L.orc_lik_export.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        L.orc_vi_run.argtypes = [C.c_int, _f64p, C.c_uint32, C.c_uint64, _f64p, _f64p, C.c_double, C.c_uint64] + [C.c_void_p] * 8
        L.orc_bootstrap_resample.argtypes = [_u64p, C.c_uint64, C.c_int32, C.c_uint64, C.c_uint64, _u32p]
        L.orc_set_num_threads.argtypes = [C.c_int]
        L.orc_dirichlet_kld.argtypes = [_f64p, C.c_uint32, C.c_uint64, _f64p, _f64p, _f64p]
        L.orc_bin_reads.argtypes = [_f64p, C.c_uint32, C.c_uint64, _f64p, C.c_void_p, _u64p, _u32p, _u64p, C.c_void_p]

    def check(self, rc: int) -> None:
        if rc != 0:
            raise RuntimeError(self.lib.orc_last_error().decode())


_libs: dict[str, _Lib] = {}


def lib(fast: bool = False) -> _Lib:
    name = "liboracle_fast.so" if fast else "liboracle.so"
    if name not in _libs:
        _libs[name] = _Lib(name)
    return _libs[name]


# ------------------------------------------------------------------------------------------------
@dataclass
class EcTable:
    n_reads: int
    n_targets: int
    hash: np.ndarray
    count: np.ndarray
    rep_read: np.ndarray
    pat_ptr: np.ndarray
    pat_targets: np.ndarray
    read_ptr: np.ndarray
    read_ids: np.ndarray
    _h: C.c_void_p

    @property
    def n_ecs(self) -> int:
        return len(self.hash)


@dataclass
class Likelihood:
    n_groups: int
    n_ecs: int
    mask: np.ndarray
    hits: np.ndarray | None
    logl: np.ndarray          # (K', N) group-major, as the reference stores it
    log_counts: np.ndarray
    lut: np.ndarray           # (K', lut_cols)
    hit_counts: np.ndarray | None   # (K, N)


@dataclass
class ViResult:
    theta: np.ndarray
    N_k: np.ndarray
    gamma: np.ndarray | None
    bound: float
    iters: int
    converged: bool
    trace_bound: np.ndarray
    trace_gnorm: np.ndarray
    trace_reset: np.ndarray
    trace_t_end: np.ndarray | None = None   # seconds since entry at the end of each iteration


def pattern_hash(targets) -> int:
    t = np.ascontiguousarray(targets, dtype=np.uint32)
    return int(lib().lib.orc_pattern_hash(t, len(t)))


def digamma(x: float) -> float:
    return float(lib().lib.orc_digamma(x))


def ldbb_scaled(k: int, n: int, a: float, b: float) -> float:
    return float(lib().lib.orc_ldbb_scaled(k, n, a, b))


def bb_parameters(n: int, q: float, e: float) -> tuple[float, float]:
    a, b = C.c_double(), C.c_double()
    lib().lib.orc_bb_parameters(n, q, e, C.byref(a), C.byref(b))
    return a.value, b.value


def _ec_from_reads(L: _Lib, reads: C.c_void_p, n_targets: int) -> EcTable:
    h = C.c_void_p()
    L.check(L.lib.orc_ec_build(reads, C.byref(h)))
    n = L.lib.orc_ec_n(h)
    nnz = L.lib.orc_ec_pat_nnz(h)
    nr = L.lib.orc_ec_n_read_ids(h)
    out = EcTable(int(L.lib.orc_reads_n(reads)), n_targets,
                  np.zeros(n, np.uint64), np.zeros(n, np.uint64), np.zeros(n, np.uint32),
                  np.zeros(n + 1, np.uint64), np.zeros(nnz, np.uint32),
                  np.zeros(n + 1, np.uint64), np.zeros(nr, np.uint32), h)
    L.lib.orc_ec_export(h, *[_opt(a) for a in (out.hash, out.count, out.rep_read, out.pat_ptr, out.pat_targets,
                                                out.read_ptr, out.read_ids)])
    return out


def ec_build_csr(n_reads: int, n_targets: int, row_ptr, targets) -> EcTable:
    L = lib()
    rp = np.ascontiguousarray(row_ptr, np.uint64)
    tg = np.ascontiguousarray(targets, np.uint32)
    if len(tg) == 0:
        tg = np.zeros(1, np.uint32)
    r = C.c_void_p()
    L.check(L.lib.orc_reads_from_csr(n_reads, n_targets, rp, tg, C.byref(r)))
    try:
        return _ec_from_reads(L, r, n_targets)
    finally:
        L.lib.orc_reads_free(r)


def read_themisto(paths: list[str], n_targets: int, merge_mode: str = "intersection"):
    """Returns (n_reads, row_ptr, targets) after strand merging, as the reference's read() would hold."""
    L = lib()
    arr = (C.c_char_p * len(paths))(*[p.encode() for p in paths])
    L.lib.orc_reads_from_files.argtypes = [C.POINTER(C.c_char_p), C.c_int, C.c_uint64, C.c_char_p, C.POINTER(C.c_void_p)]
    r = C.c_void_p()
    L.check(L.lib.orc_reads_from_files(arr, len(paths), n_targets, merge_mode.encode(), C.byref(r)))
    try:
        n = int(L.lib.orc_reads_n(r))
        nnz = int(L.lib.orc_reads_nnz(r))
        rp = np.zeros(n + 1, np.uint64)
        tg = np.zeros(max(nnz, 1), np.uint32)
        L.lib.orc_reads_csr(r, rp, tg)
        return n, rp, tg[:nnz]
    finally:
        L.lib.orc_reads_free(r)


def _read_grouping(libobj, prefix: str, path: str):
    h = C.c_void_p()
    fn = getattr(libobj, prefix + "_grouping_read")
    fn.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    if fn(path.encode(), C.byref(h)) != 0:
        err = getattr(libobj, prefix + "_last_error")
        err.restype = C.c_char_p
        raise RuntimeError(err().decode())
    ng = getattr(libobj, prefix + "_grouping_n_groups")
    nt = getattr(libobj, prefix + "_grouping_n_targets")
    nm = getattr(libobj, prefix + "_grouping_name")
    ex = getattr(libobj, prefix + "_grouping_export")
    fr = getattr(libobj, prefix + "_grouping_free")
    ng.restype, ng.argtypes = C.c_uint32, [C.c_void_p]
    nt.restype, nt.argtypes = C.c_uint64, [C.c_void_p]
    nm.restype, nm.argtypes = C.c_char_p, [C.c_void_p, C.c_uint32]
    ex.argtypes = [C.c_void_p, _u32p, _u64p]
    fr.argtypes = [C.c_void_p]
    K, T = int(ng(h)), int(nt(h))
    got = np.zeros(T, np.uint32)
    sizes = np.zeros(K, np.uint64)
    ex(h, got, sizes)
    names = [nm(h, g).decode() for g in range(K)]
    fr(h)
    return names, sizes, got


def read_grouping(path: str):
    """(names, sizes, group_of_target) from the oracle's restatement."""
    return _read_grouping(lib().lib, "orc", path)


def ref_grouping_available() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libmsweep_ref_grouping.so"))


def ref_read_grouping(path: str):
    """(names, sizes, group_of_target) from the REAL reference (src/Reference.cpp + src/Grouping.cpp)."""
    return _read_grouping(C.CDLL(os.path.join(_HERE, "_ref", "libmsweep_ref_grouping.so")), "ref", path)


def lik_build(ec: EcTable, group_of_target, sizes, q=0.65, e=0.01, zero_inflation=0.01, min_hits=0,
              keep_hit_counts=False) -> Likelihood:
    L = lib()
    got = np.ascontiguousarray(group_of_target, np.uint32)
    sz = np.ascontiguousarray(sizes, np.uint64)
    h = C.c_void_p()
    L.check(L.lib.orc_lik_build(ec._h, got, len(got), len(sz), sz, q, e, zero_inflation, min_hits,
                                int(keep_hit_counts), C.byref(h)))
    K, N, cols = int(L.lib.orc_lik_n_groups(h)), int(L.lib.orc_lik_n_ecs(h)), int(L.lib.orc_lik_lut_cols(h))
    out = Likelihood(K, N, np.zeros(len(sz), np.uint8), np.zeros(len(sz), np.uint64) if min_hits > 0 else None,
                     np.zeros((K, N)), np.zeros(N), np.zeros((K, cols)),
                     np.zeros((len(sz), N), np.uint32) if keep_hit_counts else None)
    L.lib.orc_lik_export(h, _opt(out.mask), _opt(out.hits), _opt(out.logl), _opt(out.log_counts), _opt(out.lut),
                         _opt(out.hit_counts))
    L.lib.orc_lik_free(h)
    return out


def vi_run(algo: str, logl, log_counts, alpha0=None, tol=1e-6, max_iters=5000, want_gamma=False,
           fast=False) -> ViResult:
    """algo: 'rcg' (rcgpar::rcg_optl_omp) or 'em' (rcgpar::em_torch, double). logl is (K, N) group-major."""
    L = lib(fast)
    logl = np.ascontiguousarray(logl, np.float64)
    K, N = logl.shape
    lc = np.ascontiguousarray(log_counts, np.float64)
    a0 = np.ones(K) if alpha0 is None else np.ascontiguousarray(alpha0, np.float64)
    theta, Nk, stats = np.zeros(K), np.zeros(K), np.zeros(3)
    gamma = np.zeros((K, N)) if want_gamma else None
    tb, tg, tr, tt = np.zeros(max_iters), np.zeros(max_iters), np.zeros(max_iters, np.uint8), np.zeros(max_iters)
    L.check(L.lib.orc_vi_run(0 if algo == "rcg" else 1, logl, K, N, lc, a0, tol, max_iters, _opt(theta), _opt(Nk),
                             _opt(gamma), _opt(stats), _opt(tb), _opt(tg), _opt(tr), _opt(tt)))
    it = int(stats[1])
    return ViResult(theta, Nk, gamma, float(stats[0]), it, bool(stats[2]), tb[:it], tg[:it], tr[:it], tt[:it])


def bootstrap_resample(ec_counts, seed: int, n_replicates: int, bootstrap_count: int = 0) -> np.ndarray:
    L = lib()
    c = np.ascontiguousarray(ec_counts, np.uint64)
    out = np.zeros((n_replicates, len(c)), np.uint32)
    L.check(L.lib.orc_bootstrap_resample(c, len(c), seed, bootstrap_count, n_replicates, out))
    return out


def dirichlet_kld(gamma, log_counts):
    """Sample::dirichlet_kld + get_rates (src/Sample.cpp:99-151): (log_KLD[K], RATE[K]) from the (K, N) log-posteriors."""
    L = lib()
    gamma = np.ascontiguousarray(gamma, np.float64)
    K, N = gamma.shape
    log_kld, rate = np.zeros(K), np.zeros(K)
    L.check(L.lib.orc_dirichlet_kld(gamma, K, N, np.ascontiguousarray(log_counts, np.float64), log_kld, rate))
    return log_kld, rate


def bin_reads(gamma, theta, want, read_ptr, read_ids) -> list[np.ndarray]:
    """The mGEMS rule (off-tree, unpinned): bins[k] = ascending read ids of the classes with gamma[k, j] >= log(theta[k])."""
    L = lib()
    gamma = np.ascontiguousarray(gamma, np.float64)
    K, N = gamma.shape
    theta = np.ascontiguousarray(theta, np.float64)
    want = np.ascontiguousarray(want, np.uint8)
    rp, ri = np.ascontiguousarray(read_ptr, np.uint64), np.ascontiguousarray(read_ids, np.uint32)
    ptr = np.zeros(K + 1, np.uint64)
    L.check(L.lib.orc_bin_reads(gamma, K, N, theta, want.ctypes.data, rp, ri, ptr, None))
    out = np.zeros(max(1, int(ptr[K])), np.uint32)
    L.check(L.lib.orc_bin_reads(gamma, K, N, theta, want.ctypes.data, rp, ri, ptr, out.ctypes.data))
    return [out[int(ptr[k]):int(ptr[k + 1])].copy() for k in range(K)]


def set_num_threads(n: int, fast: bool = False) -> None:
    lib(fast).lib.orc_set_num_threads(n)


def num_threads(fast: bool = False) -> int:
    return int(lib(fast).lib.orc_num_threads())
