#!/usr/bin/env python
"""bench.py — VI throughput of the B200 backend on BASELINE.json's headline config.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (config.workload): the per-GPU shard of config 3 — 1.25e7 equivalence classes x 2,000 lineages
(1e8 classes at 8 GPUs), EM/VB one-pass sweep.  The dense matrix of that config cannot exist in fp64
anywhere on the box (1.6 TB), so the default storage is the fp32 linear-domain likelihood with fp64
accumulation across classes (`--storage f64 --ecs-per-gpu 6000000` runs the fp64 series).  A "step" is
one VI iteration = one fused pass over the shard (+ the all-reduce of K+3 doubles and the control step).
Inputs are synthetic (msweep_b200/synth.py), generated on the host, far larger than L2.

Prints ONE JSON line (rank 0).  `value` is whole-job throughput with the likelihood resident in HBM,
in EC-iterations/s (classes processed per second summed over GPUs: additive, so that weak-scaling
efficiency can be computed from the per-N values); `vi_iters_per_s` is the same thing per job.
`e2e` goes through the C ABI from HOST buffers: H2D of the pseudoalignment, EC build, likelihood build,
K iterations, D2H of the abundances, all inside the timed region.

Before the timed region every run re-does one seeded small estimate (EC build -> LL_WOR21 -> RCG, EM, --min-hits)
through the same contexts — EC-sharded and hash-partitioned when N > 1 — and compares it with the frozen oracle
answers in tests/golden/bench_check.npz: `check.parity` (`check.multi_gpu_parity` at N > 1).
`extras` carries the other BASELINE configs (c1, c2, c4, c5) and config 3 in full on the lossless sparse storage
(1e8 x 2000 on ONE GPU; the same job strong-scaled at N > 1).

Nothing in the default arm executes oracle/ except the declared CPU-baseline legs (cpu_baseline, extras.c1.cpu_baseline).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_GROUPS = 2000
GROUP_SIZE = 16
FALLBACK_PEAK_GBS = 6650.0      # /opt/skills/guides/B200_PROFILING.md, used only when MEASURED_PEAKS.json is absent
METRIC = "VI throughput (EC-iterations/s)"
UNIT = "EC-iter/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ecs-per-gpu", type=int, default=12_500_000)
    ap.add_argument("--storage", default="f32", choices=["f32", "f64", "sparse"])
    ap.add_argument("--algo", default="em", choices=["em", "rcg"])
    ap.add_argument("--cpu-sample-ecs", type=int, default=200_000, help="classes of the cpu_baseline leg of the default arm")
    ap.add_argument("--ref-sample-ecs", type=int, default=1_000_000, help="classes per step of --impl reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-check", action="store_true", help="skip the seeded parity check before the timed region")
    ap.add_argument("--extras", default="auto",
                    help="auto | none | comma list of em_f64,rcg_f64,sparse,c1,c2,c4,c5 (auto: all at 1 GPU; sparse,c5 at N > 1)")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--sparse-ecs", type=int, default=100_000_000, help="classes of extras.sparse_c3_full (whole job)")
    ap.add_argument("--c4-patterns", type=int, default=50_000_000)
    ap.add_argument("--c5-replicates", type=int, default=100)
    return ap.parse_args()


def peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return FALLBACK_PEAK_GBS, "fallback (B200_PROFILING.md)"


def workload_name(a):
    st = {"f32": "fp32-stored linear likelihood, fp64 accumulation", "f64": "fp64 likelihood",
          "sparse": "lossless sparse fp64 likelihood (log(zero_inflation) once per class + its group hits)"}[a.storage]
    al = "EM/VB one-pass sweep" if a.algo == "em" else "RCG two-sweep iteration"
    return (f"config 3 shard: {a.ecs_per_gpu:.3g} ECs x {N_GROUPS} lineages per GPU ({a.ecs_per_gpu * a.gpus:.3g} ECs in the job), "
            f"{al}, {st}")


def config_of(a):
    """Identical in both arms (the driver compares them)."""
    return {"workload": workload_name(a), "n_groups": N_GROUPS, "group_size": GROUP_SIZE, "ecs_per_gpu": a.ecs_per_gpu,
            "ecs_job": a.ecs_per_gpu * a.gpus, "algo": a.algo, "storage": a.storage,
            "l2": "inputs larger than L2 (shard >> 126 MB), no flush needed", "parallelism": f"ec-shard x{a.gpus}"}


def host_threads() -> int:
    """Threads the CPU legs use: every core this process may run on — NOT the inherited OMP_NUM_THREADS (torchrun sets
    it to 1 in its children)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (recipe's clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(device_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------
# CPU legs: the oracle (a port of the reference's CPU path, reference Release flags) on a bounded sample
# ---------------------------------------------------------------------------------------------------
def cpu_vi_sample(algo: str, sample_ecs: int, steps: int, warmup: int, seed: int = 20231019):
    """Runs warmup + steps optimiser iterations of the oracle on `sample_ecs` patterns of the bench workload and times
    the last `steps`.  Returns (classes, seconds per iteration, threads, seconds of the timed region)."""
    from msweep_b200 import synth
    from oracle import pyoracle as orc
    threads = host_threads()
    orc.set_num_threads(threads, fast=True)
    orc.set_num_threads(threads, fast=False)
    wl = synth.generate_ec_patterns(sample_ecs, N_GROUPS, GROUP_SIZE, seed=seed)
    ec = orc.ec_build_csr(wl.n_reads, wl.n_targets, wl.row_ptr, wl.targets)
    lik = orc.lik_build(ec, wl.group_of_target, wl.group_sizes)
    del wl
    r = orc.vi_run(algo, lik.logl, lik.log_counts, tol=0.0 if algo == "em" else -1e300, max_iters=warmup + steps, fast=True)
    t = r.trace_t_end
    assert len(t) == warmup + steps, "the oracle stopped early"
    timed = float(t[-1] - (t[warmup - 1] if warmup > 0 else 0.0))
    return ec.n_ecs, timed / steps, orc.num_threads(fast=True), timed


def ref_sample_size(want: int) -> int:
    """The oracle holds logl and gamma (K x N doubles each) plus the build's scratch: ~3.5 matrices at the peak."""
    try:
        with open("/proc/meminfo") as f:
            avail_kb = next(int(line.split()[1]) for line in f if line.startswith("MemAvailable"))
        fit = int(avail_kb * 1024 * 0.6 / (N_GROUPS * 8 * 3.5))
        return max(20_000, min(want, fit))
    except (OSError, StopIteration):
        return want


def cpu_baseline(a, sample_ecs: int, steps: int, warmup: int):
    n, sec, threads, timed = cpu_vi_sample(a.algo, sample_ecs, steps, warmup)
    return {
        "value": n / sec, "unit": UNIT, "vi_iters_per_s_on_the_sample": 1.0 / sec, "cores": threads, "kind": "port",
        "sample": (f"oracle ({a.algo}, fp64, OpenMP x{threads}, -O3 -ffast-math) on {n} ECs x {N_GROUPS} lineages of the bench "
                   f"workload, {steps} timed iterations at {sec:.3f} s each after {warmup} warm-up; EC-iter/s is a rate: "
                   f"nothing is extrapolated"),
        "sample_ecs": n, "sample_seconds": timed,
    }


def run_reference(a):
    """--impl reference: the reference's own CPU implementation of the path on the host cores.  The reference cannot be
    built here (eight un-vendored dependencies), so this is the oracle port.  A step = one VI iteration
    over a bounded sample of the workload; the line's value is the measured rate in the bench's unit."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    t0 = time.time()
    sample = ref_sample_size(a.ref_sample_ecs)
    n, sec, threads, timed = cpu_vi_sample(a.algo, sample, a.steps, a.warmup)
    value = n / sec
    cb = {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
          "sample": (f"oracle ({a.algo}, fp64, OpenMP x{threads}, -O3 -ffast-math): each step is one VI iteration over {n} ECs x "
                     f"{N_GROUPS} lineages of the bench workload ({n * N_GROUPS * 8 / 1e9:.1f} GB of fp64 log-likelihoods); "
                     f"{a.steps} timed steps after {a.warmup} warm-up; no extrapolation (EC-iter/s is a rate)"),
          "sample_ecs": n}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * sec,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_of(a), "cpu_baseline": cb,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "extrapolated": False, "timed_region_s": timed, "wall_s": time.time() - t0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# seeded parity check against the frozen oracle answers (tests/golden/bench_check.npz)
# ---------------------------------------------------------------------------------------------------
def parity_check(M, dist, ctx, rank, world):
    from msweep_b200 import synth
    g = np.load(os.path.join(ROOT, "tests", "golden", "bench_check.npz"))
    case = eval(str(g["case"]))                                         # the dict literal make_bench_check.py wrote
    wl = synth.generate(**case)
    out = {"case": f"{case['n_reads']} reads x {case['n_targets']} refs / {case['n_groups']} lineages (seed {case['seed']})",
           "against": "tests/golden/bench_check.npz (oracle answers frozen by tests/golden/make_bench_check.py)", "world": world}
    theta_err, bound_rel, iters_ok, ecs_ok = 0.0, 0.0, True, True

    def compare(tag, res, key):
        nonlocal theta_err, bound_rel, iters_ok
        te = float(np.max(np.abs(res.theta - g[key + "_theta"])))
        br = float(abs(res.bound - float(g[key + "_bound"])) / abs(float(g[key + "_bound"])))
        ok = int(res.iters) == int(g[key + "_iters"]) and int(res.resets) == int(g[key + "_resets"])
        out[tag] = {"theta_maxabs": te, "bound_rel": br, "iters": int(res.iters), "iters_equal": ok}
        theta_err, bound_rel, iters_ok = max(theta_err, te), max(bound_rel, br), iters_ok and ok

    # (1) replicated class table, contiguous EC shards
    aln = M.Alignment(ctx, wl.n_reads, wl.n_targets, wl.row_ptr, wl.targets)
    ecs_ok = ecs_ok and aln.n_ecs == int(g["n_ecs"]) and aln.n_aligned == int(g["n_aligned"])
    if rank == 0:
        e = aln.export()
        ecs_ok = ecs_ok and int(np.bitwise_xor.reduce(e.hash)) == int(g["hash_xor"])
        ecs_ok = ecs_ok and int((e.count.astype(object) * np.arange(1, aln.n_ecs + 1).astype(object)).sum()) % (1 << 64) == int(g["count_dot"])
    lik = M.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes)
    compare("rcg", lik.vi_run(M.ALGO_RCG), "rcg")
    compare("em", lik.vi_run(M.ALGO_EM), "em")
    lik_mh = M.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes, min_hits=int(g["min_hits"]))
    mask, hits = lik_mh.mask(want_hits=True)
    out["min_hits_mask_equal"] = bool(np.array_equal(mask, g["mask_mh"]) and np.array_equal(hits, g["hits_mh"]))
    compare("rcg_min_hits", lik_mh.vi_run(M.ALGO_RCG), "rcg_mh")
    sp = M.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes, storage=M.STORE_SPARSE)
    compare("em_sparse", sp.vi_run(M.ALGO_EM), "em")
    compare("rcg_sparse", sp.vi_run(M.ALGO_RCG), "rcg")
    for x in (lik, lik_mh, sp):
        x.close()
    aln.close()

    # (2) N > 1: hash-partitioned reads (rank r owns the r-th range of the pattern hash), every rank builds its own classes
    if world > 1:
        rp = wl.row_ptr.astype(np.int64)
        lens = np.diff(rp)
        h = np.array([M.pattern_hash(wl.targets[rp[i]:rp[i + 1]]) for i in range(wl.n_reads)], np.uint64)
        owner = np.array([(int(x) * world) >> 64 for x in h], np.int64)
        mine = np.nonzero((owner == rank) & (lens > 0))[0]
        my_ptr = np.zeros(len(mine) + 1, np.uint64)
        my_ptr[1:] = np.cumsum(lens[mine])
        my_tg = np.concatenate([wl.targets[rp[i]:rp[i + 1]] for i in mine]) if len(mine) else np.zeros(0, np.uint32)
        aln_p = M.Alignment(ctx, len(mine), wl.n_targets, my_ptr, my_tg, partitioned=True)
        lik_p = M.Likelihood.build(ctx, aln_p, wl.group_of_target, wl.group_sizes)
        ecs_ok = ecs_ok and int(dist.reduce_sum(aln_p.n_ecs)) == int(g["n_ecs"]) == lik_p.n_ecs_total
        compare("rcg_hash_partitioned", lik_p.vi_run(M.ALGO_RCG), "rcg")
        compare("em_hash_partitioned", lik_p.vi_run(M.ALGO_EM), "em")
        sp_p = M.Likelihood.build(ctx, aln_p, wl.group_of_target, wl.group_sizes, storage=M.STORE_SPARSE)
        compare("rcg_sparse_hash_partitioned", sp_p.vi_run(M.ALGO_RCG), "rcg")
        sp_p.close(); lik_p.close(); aln_p.close()
    ecs_ok = bool(dist.reduce_sum(0 if ecs_ok else 1) == 0)
    out.update({"theta_maxabs": theta_err, "bound_rel": bound_rel, "iters_equal": bool(iters_ok), "n_ecs_equal": ecs_ok,
                "ok": bool(theta_err < 1e-6 and bound_rel < 1e-9 and iters_ok and ecs_ok and out["min_hits_mask_equal"]),
                "tolerances": "theta 1e-6 absolute, ELBO 1e-9 relative, identical iteration and restart counts, integers exact"})
    return out


# ---------------------------------------------------------------------------------------------------
# timed sweep series on one likelihood (resident): W warm-up + K timed iterations, CUDA events on the library's stream
# ---------------------------------------------------------------------------------------------------
def timed_series(torch, M, dist, stream, lik, algo, steps, warmup):
    """Returns dict(ms_total over ranks, pass_ms per launch, passes per iteration, bytes per iteration, launches, result)."""
    rcg = algo == M.ALGO_RCG
    sess = lik.vi_begin(algo, tol=-1e300 if rcg else 0.0, max_iters=10 ** 9, time_kernels=True)
    sess.step(warmup)
    st0 = sess.poll()
    done0 = st0.iters                                   # (RCG on several GPUs: a restart pauses the queue until a poll)
    torch.cuda.synchronize()
    dist.barrier()
    launches0 = M.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    sess.step(steps)
    st1 = sess.poll() if rcg and lik.ctx.world_size > 1 else None
    while st1 is not None and st1.iters - done0 < steps:
        sess.step(steps - (st1.iters - done0))
        st1 = sess.poll()
    ev1.record(stream)
    torch.cuda.synchronize()
    dist.barrier()
    launches = M.launch_count() - launches0
    ms_total = dist.reduce_max(ev0.elapsed_time(ev1))
    st1 = sess.poll()
    res = sess.finish()
    n_it = st1.iters - st0.iters
    assert n_it == steps, f"the timed region must contain exactly --steps iterations (got {n_it})"
    n_pass = max(1, st1.pass_launches - st0.pass_launches)
    return {"ms_total": ms_total, "pass_ms": (st1.pass_ms_sum - st0.pass_ms_sum) / n_pass, "passes_per_iter": n_pass / steps,
            "bytes_per_iter": st1.pass_bytes, "launches": launches, "res": res}


def series_summary(s, steps, n_ecs, n_groups, peak):
    kms = s["pass_ms"] * s["passes_per_iter"]
    gbs = s["bytes_per_iter"] / (kms * 1e-3) / 1e9
    ms = s["ms_total"] / steps
    return {"ecs": int(n_ecs), "n_groups": int(n_groups), "steps": steps, "ms_per_step": ms, "vi_iters_per_s": 1e3 / ms,
            "pass_kernels_ms_per_step": kms, "achieved_gbs": gbs, "frac_of_measured_peak": gbs / peak,
            "frac_of_nominal_8TBs": gbs / 8000.0, "bytes_per_step": int(s["bytes_per_iter"]),
            "launches_per_step": s["launches"] / steps, "kernel_share_of_step": kms / ms}


# ---------------------------------------------------------------------------------------------------
# extras: the other BASELINE configs
# ---------------------------------------------------------------------------------------------------
def extra_c1(a):
    """Config 1 end to end through the binaries: 1e6 paired reads (Themisto text) x 3000 refs / 50 lineages, RCG.
    mSWEEP_b200 against the oracle's CLI (rcgcpu restatement, -t 8) on the same files."""
    from msweep_b200 import synth
    d = tempfile.mkdtemp(prefix="mswb_c1_")
    t0 = time.time()
    wl = synth.generate(1_000_000, 3000, 50, n_present=5, n_templates=2000, p_noise=0.02, seed=20231017)
    paths = synth.write_themisto(os.path.join(d, "aln"), wl, paired=True, shuffle_frac=0.01)
    gfile = os.path.join(d, "grouping.txt")
    synth.write_grouping(gfile, wl)
    t_gen = time.time() - t0
    common = ["--themisto-1", paths[0], "--themisto-2", paths[1], "-i", gfile, "-t", "8"]
    env = dict(os.environ, OMP_NUM_THREADS="8")
    t0 = time.time()
    r = subprocess.run([os.path.join(ROOT, "msweep_b200", "bin", "mSWEEP_b200"), *common, "-o", os.path.join(d, "ours"), "--print-timings"],
                       capture_output=True, text=True, env=env)
    t_ours = time.time() - t0
    if r.returncode != 0:
        return {"error": r.stderr[-400:]}
    stages = json.loads(r.stderr.strip().splitlines()[-1])
    t0 = time.time()
    r = subprocess.run([os.path.join(ROOT, "oracle", "msweep_oracle"), *common, "-o", os.path.join(d, "ref"), "--algorithm", "rcgcpu", "--print-timings"],
                       capture_output=True, text=True, env=env)
    t_ref = time.time() - t0
    if r.returncode != 0:
        return {"error": "oracle CLI: " + r.stderr[-400:]}
    ref_stages = json.loads(r.stderr.strip().splitlines()[-1])

    def vals(p):
        lines = open(p).read().splitlines()
        rows = [l.split("\t") for l in lines if not l.startswith("#")]
        return [l for l in lines if l.startswith("#")][1:], [x[0] for x in rows], np.array([float(x[1]) for x in rows])

    h1, n1, v1 = vals(os.path.join(d, "ours_abundances.txt"))
    h2, n2, v2 = vals(os.path.join(d, "ref_abundances.txt"))
    work = stages["parse_s"] + stages["ec_build_s"] + stages["likelihood_s"] + stages["optimiser_s"] + stages["write_s"]
    out = {"config": "1: 1e6 paired reads x 3000 refs / 50 lineages, Themisto text in, abundances out, RCG, -t 8",
           "input_mb": sum(os.path.getsize(p) for p in paths) / 1e6, "generate_s": round(t_gen, 1),
           "msweep_b200": {"process_wall_s": round(t_ours, 3), "work_s": round(work, 4), "stages": stages,
                           "us_per_rcg_iteration": 1e6 * stages["optimiser_s"] / max(1, stages["iters"])},
           "cpu_baseline": {"kind": "port", "what": "oracle CLI (rcgcpu restatement), -t 8", "process_wall_s": round(t_ref, 3), "stages": ref_stages},
           "same_header_lines": h1 == h2, "same_group_order": n1 == n2, "max_abs_theta_diff_as_printed": float(np.max(np.abs(v1 - v2))),
           "iters_equal": int(stages["iters"]) == int(ref_stages.get("optimiser_iters", -1))}
    for p in os.listdir(d):
        os.unlink(os.path.join(d, p))
    os.rmdir(d)
    return out


def run_to_convergence(M, ctx, lik, algo, max_iters=5000):
    t0 = time.perf_counter()
    r = lik.vi_run(algo, tol=1e-6, max_iters=max_iters, time_kernels=True)
    ctx.sync()
    dt = time.perf_counter() - t0
    kms = r.pass_ms_sum / max(1, r.iters)
    return r, {"seconds": round(dt, 4), "iters": int(r.iters), "converged": bool(r.converged), "resets": int(r.resets), "ms_per_iter_wall": dt / max(1, r.iters) * 1e3,
               "pass_kernels_ms_per_iter": kms, "achieved_gbs": r.pass_bytes / (kms * 1e-3) / 1e9 if kms > 0 else None,
               "bound": float(r.bound), "theta_sum": float(r.theta.sum())}


def extra_c2_c5(a, torch, M, dist, ctx, stream, rank, world, want_c2, want_c5, peak):
    """Config 2 / 5 shape: 1e7 reads -> ~1e6 ECs x 60,000 refs in 1,000 lineages, fp64.  c2: RCG and EM to convergence on one
    GPU.  c5: --iters bootstrap replicates, replicate r on rank r % N (the likelihood is replicated: replicas only)."""
    from msweep_b200 import synth
    out = {}
    t0 = time.time()
    wl = synth.generate_ec_patterns(1_000_000, 1000, 60, n_present=20, seed=55, dup_factor=9.0)
    t_gen = time.time() - t0
    solo = M.Context(ctx.device, 0, 1, None, cuda_stream=stream.cuda_stream) if world > 1 else ctx   # replicas: no collective
    with torch.cuda.stream(stream):
        t0 = time.perf_counter()
        aln = M.Alignment(solo, wl.n_reads, wl.n_targets, wl.row_ptr, wl.targets)
        solo.sync(); t_ec = time.perf_counter() - t0
        t0 = time.perf_counter()
        lik = M.Likelihood.build(solo, aln, wl.group_of_target, wl.group_sizes)
        solo.sync(); t_lik = time.perf_counter() - t0
        shape = {"reads": int(wl.n_reads), "aligned": int(aln.n_aligned), "ecs": int(aln.n_ecs), "refs": int(wl.n_targets), "lineages": 1000,
                 "generate_s": round(t_gen, 1), "ec_build_s": round(t_ec, 4), "likelihood_s": round(t_lik, 4)}
        if want_c2:
            r_rcg, s_rcg = run_to_convergence(M, solo, lik, M.ALGO_RCG)
            r_em, s_em = run_to_convergence(M, solo, lik, M.ALGO_EM)
            for s in (s_rcg, s_em):
                if s["achieved_gbs"]:
                    s["frac_of_measured_peak"] = s["achieved_gbs"] / peak
            kept = wl.truth > 0
            t0 = time.perf_counter()
            sp = M.Likelihood.build(solo, aln, wl.group_of_target, wl.group_sizes, storage=M.STORE_SPARSE)
            solo.sync(); t_lik_sp = time.perf_counter() - t0
            r_rcg_sp, s_rcg_sp = run_to_convergence(M, solo, sp, M.ALGO_RCG)
            r_em_sp, s_em_sp = run_to_convergence(M, solo, sp, M.ALGO_EM)
            sp.close()
            for s_, r_, ref_ in ((s_rcg_sp, r_rcg_sp, r_rcg), (s_em_sp, r_em_sp, r_em)):
                s_["frac_of_measured_peak_own_bytes"] = s_["achieved_gbs"] / peak if s_["achieved_gbs"] else None
                s_["max_abs_theta_diff_vs_dense"] = float(np.max(np.abs(r_.theta - ref_.theta)))
                s_["iters_equal_dense"] = int(r_.iters) == int(ref_.iters)
            out["c2"] = {"config": "2: 1e7 reads / ~1e6 ECs x 60,000 refs in 1,000 lineages, fp64 VI on 1 GPU", **shape, "rcg": s_rcg, "em": s_em,
                         "sparse_likelihood_s": round(t_lik_sp, 4), "rcg_sparse": s_rcg_sp, "em_sparse": s_em_sp,
                         "e2e_seconds_from_host_csr_rcg_sparse": round(t_ec + t_lik_sp + s_rcg_sp["seconds"], 4),
                         "max_abs_theta_diff_em_vs_rcg": float(np.max(np.abs(r_rcg.theta - r_em.theta))),
                         "max_abs_err_vs_generating_theta": float(np.max(np.abs(r_rcg.theta - wl.truth))),
                         "present_lineages_recovered": int(np.sum(r_rcg.theta[kept] > 1e-4)), "present_lineages": int(kept.sum()),
                         "e2e_seconds_from_host_csr_rcg": round(t_ec + t_lik + s_rcg["seconds"], 4)}
        if want_c5:
            B = a.c5_replicates
            # the share of this rank when B replicates are spread over 8 GPUs — with N < 8 the job is weak-scaled (N/8 of config 5)
            n_rep_job = B if world >= 8 else int(np.ceil(B / 8)) * world
            res = {}
            sparse = M.Likelihood.build(solo, aln, wl.group_of_target, wl.group_sizes, storage=M.STORE_SPARSE)
            for name, algo, L in (("rcg", M.ALGO_RCG, lik), ("em_batched", M.ALGO_EM, lik), ("em_sparse", M.ALGO_EM, sparse),
                                  ("rcg_sparse", M.ALGO_RCG, sparse)):
                runs = []
                for _ in range(3 if L is sparse else 1):          # the sub-second jobs are timed three times (host jitter): all runs reported
                    dist.barrier()
                    t0 = time.perf_counter()
                    thetas, iters = L.bootstrap_run(n_rep_job, seed=11, algo=algo, replica_rank=rank, replica_world=world)
                    solo.sync()
                    runs.append(round(dist.reduce_max(time.perf_counter() - t0), 3))
                sec = float(np.median(runs))
                mine = [i for i in range(n_rep_job) if i % world == rank]
                res[name] = {"replicates_in_job": n_rep_job, "replicates_on_rank0": len(mine), "seconds": round(sec, 3), "seconds_runs": runs,
                             "replicates_per_s": n_rep_job / sec, "mean_iters": float(np.mean([iters[i] for i in mine])),
                             "theta_sum_min": float(np.min(thetas[mine].sum(axis=1))), "theta_sum_max": float(np.max(thetas[mine].sum(axis=1)))}
                if name == "rcg":
                    th_rcg = thetas[mine]
                else:
                    res[name]["max_abs_theta_diff_vs_rcg_replicates"] = float(np.max(np.abs(thetas[mine] - th_rcg)))
            sparse.close()
            res["how"] = {"rcg": "one replicate after the other, each a cold-start RCG run (the reference's default algorithm)",
                          "em_batched": "dense fp64: all count vectors of the rank resampled first, then ONE sweep of the matrix per iteration serves every replicate still running (em_lin_batch_kernel)",
                          "em_sparse": "lossless sparse storage: one replicate after the other, each pass reads ~90 B per class instead of 8 KB",
                          "rcg_sparse": "RCG on the sparse storage (separable state off the hits): one replicate after the other, the reference's default algorithm"}
            out["c5"] = {"config": f"5: --iters {B} bootstrap on 1e7 reads x 1,000 lineages, replicates spread over 8 GPUs "
                                   f"(this run: {n_rep_job} replicates on {world} GPU(s), replicas only, exact std::mt19937_64 resampling)",
                         **shape, **res}
        lik.close(); aln.close()
    if solo is not ctx:
        solo.close()
    return out


def extra_c4(a, torch, M, ctx, stream, peak):
    """Config 4: 5e7 patterns x 10,000 lineages, most of them empty, --min-hits 1 (mask + compaction), then VI on K' x N."""
    from msweep_b200 import synth
    t0 = time.time()
    wl = synth.generate_ec_patterns(a.c4_patterns, 10_000, 6, n_present=50, n_pool=100, seed=20231021)
    t_gen = time.time() - t0
    hit_groups = np.zeros(10_000, bool)
    hit_groups[np.unique(wl.group_of_target[np.unique(wl.targets)])] = True       # host tally: groups that receive any hit
    with torch.cuda.stream(stream):
        t0 = time.perf_counter(); aln = M.Alignment(ctx, wl.n_reads, wl.n_targets, wl.row_ptr, wl.targets); ctx.sync(); t_ec = time.perf_counter() - t0
        t0 = time.perf_counter(); lik = M.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes, min_hits=1); ctx.sync(); t_lik = time.perf_counter() - t0
        mask, hits = lik.mask(want_hits=True)
        out = {"config": f"4: {a.c4_patterns:.3g} patterns x 10,000 lineages (most empty), --min-hits 1, 1 GPU", "generate_s": round(t_gen, 1),
               "ecs": int(aln.n_ecs), "ec_build_s": round(t_ec, 4), "likelihood_mask_compaction_s": round(t_lik, 4), "groups_kept": int(lik.n_groups),
               "mask_matches_host_tally": bool(np.array_equal(mask.astype(bool), hit_groups)),
               "zeros_last_ordering": "pruned lineages are written after the kept ones (tests/test_gpu_cli.py::test_min_hits_orders_pruned_groups_last)"}
        kept = np.flatnonzero(mask)
        r_rcg, out["rcg"] = run_to_convergence(M, ctx, lik, M.ALGO_RCG)
        r_em, out["em"] = run_to_convergence(M, ctx, lik, M.ALGO_EM)
        for s in (out["rcg"], out["em"]):
            if s["achieved_gbs"]:
                s["frac_of_measured_peak"] = s["achieved_gbs"] / peak
        out["max_abs_theta_diff_em_vs_rcg"] = float(np.max(np.abs(r_rcg.theta - r_em.theta)))
        out["max_abs_err_vs_generating_theta"] = float(np.max(np.abs(r_rcg.theta - wl.truth[kept])))
        lik.close()
        t0 = time.perf_counter(); sp = M.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes, min_hits=1, storage=M.STORE_SPARSE); ctx.sync()
        out["sparse_likelihood_mask_compaction_s"] = round(time.perf_counter() - t0, 4)
        r_sp, out["rcg_sparse"] = run_to_convergence(M, ctx, sp, M.ALGO_RCG)
        out["rcg_sparse"]["max_abs_theta_diff_vs_dense"] = float(np.max(np.abs(r_sp.theta - r_rcg.theta)))
        out["rcg_sparse"]["iters_equal_dense"] = int(r_sp.iters) == int(r_rcg.iters)
        r_sp, out["em_sparse"] = run_to_convergence(M, ctx, sp, M.ALGO_EM)
        out["em_sparse"]["max_abs_theta_diff_vs_dense"] = float(np.max(np.abs(r_sp.theta - r_em.theta)))
        sp.close(); aln.close()
    return out


def extra_sparse(a, torch, M, dist, ctx, stream, rank, world, wl_main, peak):
    """Config 3 IN FULL — 1e8 classes x 2000 lineages — on the lossless sparse storage: one GPU holds it all; at N > 1 the same
    job is strong-scaled (EC shards, the same all-reduce).  Resident series + end to end from host buffers."""
    from msweep_b200 import synth
    n_local = a.sparse_ecs // world
    t0 = time.time()
    wl = wl_main if n_local == wl_main.n_reads else synth.generate_ec_patterns(n_local, N_GROUPS, GROUP_SIZE, seed=20231019 + rank)
    t_gen = time.time() - t0
    steps, warmup = 50, 3
    with torch.cuda.stream(stream):
        aln = M.Alignment(ctx, wl.n_reads, wl.n_targets, wl.row_ptr, wl.targets, partitioned=world > 1)
        lik = M.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes, storage=M.STORE_SPARSE)
        n_job = dist.reduce_sum(lik.n_ecs)
        s = timed_series(torch, M, dist, stream, lik, M.ALGO_EM, steps, warmup)
        out = series_summary(s, steps, lik.n_ecs, N_GROUPS, peak)
        s_rcg = timed_series(torch, M, dist, stream, lik, M.ALGO_RCG, 20, warmup)
        out["rcg"] = series_summary(s_rcg, 20, lik.n_ecs, N_GROUPS, peak)
        out["rcg"]["what"] = ("RCG (the reference's default optimiser) on the same job: dense state would be 3.2 TB; the sparse form keeps "
                              "two K-vectors, two N-vectors and the hits")
        lik.close(); aln.close()
        if world > 1:
            # A/B of the per-pass collective on this very job: the library's one-shot exchange over NVLink peer memory inside
            # its control kernel (csrc/peer.cuh) against ncclAllReduce + a control kernel (a second context with MSWB_PEER=0)
            out["collective"] = "peer-memory one-shot exchange (in-kernel, rank-ordered sum)" if ctx.peer_active else "ncclAllReduce"
            if ctx.peer_active:
                try:
                    os.environ["MSWB_PEER"] = "0"
                    nid = dist.broadcast_bytes(M.nccl_unique_id() if rank == 0 else None, M.NCCL_ID_BYTES)
                    ctx2 = M.Context(ctx.device, rank, world, nid, cuda_stream=stream.cuda_stream)
                finally:
                    os.environ.pop("MSWB_PEER", None)
                aln2 = M.Alignment(ctx2, wl.n_reads, wl.n_targets, wl.row_ptr, wl.targets, partitioned=True)
                lik2 = M.Likelihood.build(ctx2, aln2, wl.group_of_target, wl.group_sizes, storage=M.STORE_SPARSE)
                s2 = timed_series(torch, M, dist, stream, lik2, M.ALGO_EM, steps, warmup)
                s2r = timed_series(torch, M, dist, stream, lik2, M.ALGO_RCG, 20, warmup)
                out["collective_ab"] = {"em_ms_per_step": {"peer": s["ms_total"] / steps, "nccl": s2["ms_total"] / steps},
                                        "rcg_ms_per_step": {"peer": s_rcg["ms_total"] / 20, "nccl": s2r["ms_total"] / 20},
                                        "em_launches_per_step": {"peer": s["launches"] / steps, "nccl": s2["launches"] / steps},
                                        "theta_maxabs_peer_vs_nccl": float(np.max(np.abs(s["res"].theta - s2["res"].theta))),
                                        "nccl_peer_active": bool(ctx2.peer_active)}
                lik2.close(); aln2.close(); ctx2.close()
        pinned = []
        if wl is not wl_main:                                   # the e2e leg copies from PINNED host memory, like the headline's
            for arr in (wl.row_ptr, wl.targets):
                if int(torch.cuda.cudart().cudaHostRegister(arr.ctypes.data, arr.nbytes, 0)) == 0:
                    pinned.append(arr)
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        aln = M.Alignment(ctx, wl.n_reads, wl.n_targets, wl.row_ptr, wl.targets, partitioned=world > 1)
        lik = M.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes, storage=M.STORE_SPARSE)
        r = lik.vi_run(M.ALGO_EM, tol=0.0, max_iters=steps, poll_every=steps)
        torch.cuda.synchronize()
        sec = dist.reduce_max(time.perf_counter() - t0)
        lik.close(); aln.close()
        for arr in pinned:
            torch.cuda.cudart().cudaHostUnregister(arr.ctypes.data)
    out.update({"config": f"3 in full: {a.sparse_ecs:.3g} ECs x {N_GROUPS} lineages, EM/VB, lossless sparse fp64 storage, "
                          f"{'one GPU' if world == 1 else f'strong-scaled over {world} GPUs (EC shards, one all-reduce of K+3 doubles per pass)'}",
                "ecs_job": int(n_job), "ecs_per_gpu": int(out.pop("ecs")), "scaling": "strong", "generate_s": round(t_gen, 1),
                "value_ec_iter_per_s": n_job * steps / (s["ms_total"] * 1e-3),
                "roofline": {"bound": "hbm", "kernel": "em_sparse_pass_kernel", "achieved": out["achieved_gbs"], "peak": peak, "unit": "GB/s",
                             "frac": out["achieved_gbs"] / peak, "bytes_per_launch": out["bytes_per_step"],
                             "note": "its own algorithmic bytes: 12 B per (class, group) hit + 32 B per class; the dense fp32 form of the same job reads 800 GB per pass"},
                "e2e": {"seconds": round(sec, 4), "vi_iters_per_s": steps / sec, "value_ec_iter_per_s": n_job * steps / sec,
                        "what": f"mswb_ec_build + mswb_lik_build(sparse) + {steps} iterations from host CSR buffers, theta back on the host",
                        "h2d_bytes": int(wl.row_ptr.nbytes + wl.targets.nbytes)},
                "check": {"theta_sum": float(r.theta.sum()), "bound": float(r.bound)}})
    return out


# ---------------------------------------------------------------------------------------------------
def main():
    a = parse_args()
    if a.impl == "reference":
        return run_reference(a)

    import torch
    import msweep_b200 as M
    from msweep_b200 import dist, synth

    rank, world, local = dist.init()
    assert world == a.gpus, f"--gpus {a.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {a.gpus}"
    torch.cuda.set_device(local)
    M.lib()                                                   # raises if the CUDA library is missing: no fallback
    stream = torch.cuda.Stream()
    nccl_id = None
    if world > 1:
        nccl_id = dist.broadcast_bytes(M.nccl_unique_id() if rank == 0 else None, M.NCCL_ID_BYTES)
    ctx = M.Context(local, rank, world, nccl_id, cuda_stream=stream.cuda_stream)
    peak, peak_src = peak_gbs()
    extras_sel = set() if (a.no_extras or a.extras == "none") else (
        ({"em_f64", "rcg_f64", "sparse", "c1", "c2", "c4", "c5"} if world == 1 else {"sparse", "c5"}) if a.extras == "auto"
        else set(a.extras.split(",")))
    if not (a.storage == "f32" and a.algo == "em"):
        extras_sel = set()                                    # side series belong to the default line only

    # ---- seeded parity check through these very contexts, before anything is timed ---------------------
    check = None
    if not a.no_check:
        with torch.cuda.stream(stream):
            check = parity_check(M, dist, ctx, rank, world)

    n_local = a.ecs_per_gpu
    storage = {"f32": M.STORE_F32, "f64": M.STORE_F64, "sparse": M.STORE_SPARSE}[a.storage]
    algo = M.ALGO_EM if a.algo == "em" else M.ALGO_RCG
    t_gen = time.time()
    wl = synth.generate_ec_patterns(n_local, N_GROUPS, GROUP_SIZE, seed=20231019 + rank)
    t_gen = time.time() - t_gen
    # the e2e leg copies its inputs from PINNED host memory (page-locked in place)
    for arr in (wl.row_ptr, wl.targets, wl.group_of_target, wl.group_sizes):
        torch.cuda.cudart().cudaHostRegister(arr.ctypes.data, arr.nbytes, 0)

    # ---- resident leg: likelihood in HBM, W warm-up + K timed iterations ----------------------------
    with torch.cuda.stream(stream):
        aln = M.Alignment(ctx, wl.n_reads, wl.n_targets, wl.row_ptr, wl.targets, partitioned=world > 1)
        lik = M.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes, storage=storage)
        n_ecs_local = lik.n_ecs
        sampler = ClockSampler(local) if rank == 0 else None
        s = timed_series(torch, M, dist, stream, lik, algo, a.steps, a.warmup)
        clocks = sampler.stop() if sampler else None
    res = s["res"]
    # per-rank kernel time: at N > 1 a step lasts as long as the SLOWEST rank's pass (the others wait for it in the all-reduce)
    pass_ms_max, pass_ms_min = dist.reduce_max(s["pass_ms"]), -dist.reduce_max(-s["pass_ms"])
    n_job = dist.reduce_sum(n_ecs_local)
    ms_per_step = s["ms_total"] / a.steps
    value = n_job * a.steps / (s["ms_total"] * 1e-3)
    bytes_per_launch = s["bytes_per_iter"] / s["passes_per_iter"]
    theta_sum = float(res.theta.sum())
    lik.close(); aln.close()

    # ---- end-to-end leg: host buffers -> abundances through the C ABI, copies inside the timed region --
    e2e = None
    if not a.no_e2e:
        ctx.trim()                                                # cold allocator: nothing parked from the resident leg
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        with torch.cuda.stream(stream):
            aln = M.Alignment(ctx, wl.n_reads, wl.n_targets, wl.row_ptr, wl.targets, partitioned=world > 1)
            t_ec = time.perf_counter()              # (mswb_ec_build returns with the table built: its class count is an output)
            lik = M.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes, storage=storage)
            t_lik = time.perf_counter()             # (enqueued, not necessarily finished: no synchronisation is added for the split)
            r2 = lik.vi_run(algo, tol=0.0 if a.algo == "em" else -1e300, max_iters=a.steps, poll_every=a.steps)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        sec = dist.reduce_max(t1 - t0)
        stages = {"ec_build_s": round(t_ec - t0, 4), "likelihood_build_call_s": round(t_lik - t_ec, 4), "vi_run_s": round(t1 - t_lik, 4)}
        assert r2.iters == a.steps
        h2d = wl.row_ptr.nbytes + wl.targets.nbytes + wl.group_of_target.nbytes + wl.group_sizes.nbytes + 8 * N_GROUPS
        e2e = {"value": n_job * a.steps / sec, "unit": UNIT, "vi_iters_per_s": a.steps / sec, "seconds": sec,
               "h2d_bytes_per_step": h2d / a.steps, "d2h_bytes_per_step": (8 * N_GROUPS + 64) / a.steps, "stages_rank0": stages,
               "what": "mswb_ec_build + mswb_lik_build + mswb_vi_run(K iterations) from pinned host CSR buffers, theta back on the host",
               "allocator": "cold: the library's device block cache was emptied (mswb_ctx_trim) before the leg, so the cudaMalloc of the matrix is inside"}
        lik.close(); aln.close()

    # ---- side series and the other BASELINE configs ---------------------------------------------------------
    extras = {}

    def guarded(name, fn):
        t0 = time.time()
        try:
            out = fn()
        except Exception as e:                                   # an extra must never cost the headline line
            out = {"error": f"{type(e).__name__}: {e}"[:500]}
        if isinstance(out, dict):
            if set(out) <= {"c2", "c5"} and out:
                for k, v in out.items():
                    extras[k] = v
            else:
                out["wall_s"] = round(time.time() - t0, 1)
                extras[name] = out

    for name, st, al, n_sub, steps in (("em_f64", M.STORE_F64, M.ALGO_EM, min(n_local, 6_000_000), 20),
                                       ("rcg_f64", M.STORE_F64, M.ALGO_RCG, min(n_local, 1_000_000), 10)):
        if name not in extras_sel:
            continue

        def run(st=st, al=al, n_sub=n_sub, steps=steps):
            rp = wl.row_ptr[:n_sub + 1]
            with torch.cuda.stream(stream):
                aln = M.Alignment(ctx, n_sub, wl.n_targets, rp, wl.targets[:int(rp[-1])])
                lik = M.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes, storage=st)
                sx = timed_series(torch, M, dist, stream, lik, al, steps, a.warmup)
                out = series_summary(sx, steps, lik.n_ecs, N_GROUPS, peak)
                lik.close(); aln.close()
            return out
        guarded(name, run)
    if "sparse" in extras_sel:
        guarded("sparse_c3_full", lambda: extra_sparse(a, torch, M, dist, ctx, stream, rank, world, wl, peak))
    for arr in (wl.row_ptr, wl.targets, wl.group_of_target, wl.group_sizes):
        torch.cuda.cudart().cudaHostUnregister(arr.ctypes.data)
    n_reads_main, nbytes_main = wl.n_reads, wl.row_ptr.nbytes + wl.targets.nbytes
    del wl
    ctx.trim()
    if "c2" in extras_sel or "c5" in extras_sel:
        guarded("c2_c5", lambda: extra_c2_c5(a, torch, M, dist, ctx, stream, rank, world, "c2" in extras_sel and world == 1, "c5" in extras_sel, peak))
    if "c4" in extras_sel and world == 1:
        guarded("c4", lambda: extra_c4(a, torch, M, ctx, stream, peak))
    if "c1" in extras_sel and world == 1 and rank == 0:
        ctx.trim()      # the binary is another process: do not leave it a device whose memory this one has parked
        guarded("c1", lambda: extra_c1(a))

    cb = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cb = cpu_baseline(a, a.cpu_sample_ecs, steps=3, warmup=1)

    if rank == 0:
        achieved = bytes_per_launch / (s["pass_ms"] * 1e-3) / 1e9
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                tj = json.load(f).get(f"{a.algo}_{a.storage}")
            if tj:   # dram bytes per element from the ncu --set full capture, scaled to this launch
                traffic = tj["dram_bytes_per_element"] * n_ecs_local * N_GROUPS
                traffic_src = "ncu --set full capture (profiles/traffic.json), bytes per element x this launch's elements; not an in-run counter"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT,
            "vi_iters_per_s": a.steps / (s["ms_total"] * 1e-3), "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 accumulation / f32 storage" if a.storage == "f32" else "f64", "data": "synthetic",
            "config": config_of(a),
            "workload_built": {"ecs_per_gpu": int(n_ecs_local), "ecs_job": int(n_job), "generator_s": round(t_gen, 1),
                               "reads_per_gpu": int(n_reads_main), "csr_bytes_per_gpu": int(nbytes_main)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "frac_of_nominal_8TBs": achieved / 8000.0, "peak_source": peak_src, "traffic": traffic, "traffic_source": traffic_src,
                         "kernel": ("em_sparse_pass_kernel" if a.storage == "sparse" else "em_lin_pass_kernel") if a.algo == "em" else "rcg_sweep_a_kernel + rcg_sweep_b_kernel",
                         "bytes_per_launch": bytes_per_launch, "kernel_ms": s["pass_ms"],
                         "kernel_ms_over_ranks": {"min": pass_ms_min, "max": pass_ms_max},
                         "kernel_share_of_step": s["pass_ms"] * s["passes_per_iter"] / ms_per_step,
                         "slowest_rank_kernel_share_of_step": pass_ms_max * s["passes_per_iter"] / ms_per_step},
            "cpu_baseline": cb,
            "e2e": e2e,
            "gpu_launches": s["launches"],
            "collective": None if world == 1 else ("peer-memory one-shot exchange inside the control kernel (csrc/peer.cuh)" if ctx.peer_active
                                                  else "ncclAllReduce + control kernel"),
            "clocks": clocks,
            "extras": extras or None,
            "check": {"theta_sum": theta_sum, "bound": res.bound, ("multi_gpu_parity" if world > 1 else "parity"): check},
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    dist.finalize()


if __name__ == "__main__":
    main()
