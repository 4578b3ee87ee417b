#!/usr/bin/env python
"""bench.py — VI throughput of the B200 backend on BASELINE.json's headline config.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (config.workload): the per-GPU shard of config 3 — 1.25e7 equivalence classes x 2,000 lineages
(1e8 classes at 8 GPUs), EM/VB one-pass sweep.  The dense matrix of that config cannot exist in fp64
anywhere on the box (1.6 TB), so the default storage is the fp32 linear-domain likelihood with fp64
accumulation across classes (`--storage f64 --ecs-per-gpu 6000000` runs the fp64 series).  A "step" is
one VI iteration = one fused pass over the shard + the all-reduce of K+1 doubles + the control kernel.
Inputs are synthetic (msweep_b200/synth.py), generated on the host, far larger than L2.

Prints ONE JSON line (rank 0).  `value` is whole-job throughput with the likelihood resident in HBM,
in EC-iterations/s (classes processed per second summed over GPUs: additive, so that weak-scaling
efficiency can be computed from the per-N values); `vi_iters_per_s` is the same thing per job.
`e2e` goes through the C ABI from HOST buffers: H2D of the pseudoalignment, EC build, likelihood build,
K iterations, D2H of the abundances, all inside the timed region.
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ecs-per-gpu", type=int, default=12_500_000)
    ap.add_argument("--storage", default="f32", choices=["f32", "f64", "sparse"])
    ap.add_argument("--algo", default="em", choices=["em", "rcg"])
    ap.add_argument("--cpu-sample-ecs", type=int, default=40_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the fp64 EM and RCG side series (N = 1 only)")
    return ap.parse_args()


def peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return FALLBACK_PEAK_GBS, "fallback (B200_PROFILING.md)"


def workload_name(a, n_local):
    st = {"f32": "fp32-stored linear likelihood, fp64 accumulation", "f64": "fp64 likelihood",
          "sparse": "lossless sparse fp64 likelihood (log(zero_inflation) once per class + its group hits)"}[a.storage]
    al = "EM/VB one-pass sweep" if a.algo == "em" else "RCG two-sweep iteration"
    return (f"config 3 shard: {n_local:.3g} ECs x {N_GROUPS} lineages per GPU ({n_local * a.gpus:.3g} ECs in the job), "
            f"{al}, {st}")


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
# CPU baseline: the oracle (a port of the reference's CPU path) on a bounded sample of the same workload
# ---------------------------------------------------------------------------------------------------
def cpu_baseline(wl, a, n_local, steps, warmup):
    """Times `steps` optimiser iterations of the oracle (reference Release flags, all host threads) on the
    first `cpu_sample_ecs` patterns of the workload and scales to the shard linearly in the class count."""
    from oracle import pyoracle as orc
    ns = min(a.cpu_sample_ecs, wl.n_reads)
    rp = wl.row_ptr[:ns + 1]
    ec = orc.ec_build_csr(ns, wl.n_targets, rp, wl.targets[:int(rp[-1])])
    lik = orc.lik_build(ec, wl.group_of_target, wl.group_sizes)
    threads = orc.num_threads(fast=True)
    r = orc.vi_run(a.algo, lik.logl, lik.log_counts, tol=0.0 if a.algo == "em" else -1e300, max_iters=warmup + steps, fast=True)
    t = r.trace_t_end
    n_done = len(t)
    w = min(warmup, n_done - 1)
    sec_per_iter = (t[-1] - (t[w - 1] if w > 0 else 0.0)) / max(1, n_done - w)
    scale = n_local / ec.n_ecs                       # cost is linear in the number of classes
    iters_per_s_shard = 1.0 / (sec_per_iter * scale)
    return {
        "value": iters_per_s_shard * n_local,      # EC-iterations/s, the line's unit
        "unit": "EC-iter/s",
        "vi_iters_per_s": iters_per_s_shard,
        "cores": threads,
        "kind": "port",
        "sample": (f"oracle ({a.algo}, fp64, OpenMP x{threads}, -O3 -ffast-math) on {ec.n_ecs} ECs x {N_GROUPS} lineages, "
                   f"{n_done - w} timed iterations at {sec_per_iter:.3f} s each, scaled linearly to {n_local} ECs"),
    }


def run_reference(a):
    """--impl reference: the reference's own CPU implementation of the path.  The reference cannot be
    built here (eight un-vendored dependencies), so this is the oracle port."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from msweep_b200 import synth
    wl = synth.generate_ec_patterns(a.cpu_sample_ecs, N_GROUPS, GROUP_SIZE, seed=20231019)
    t0 = time.time()
    cb = cpu_baseline(wl, a, a.ecs_per_gpu, a.steps, a.warmup)
    n_job = a.ecs_per_gpu * a.gpus
    value = cb["value"]      # EC-iterations/s of the host cores; the whole job's classes go through the same host
    line = {
        "impl": "reference", "metric": "VI throughput (EC-iterations/s)", "value": value, "unit": "EC-iter/s",
        "vi_iters_per_s": value / n_job, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3 * n_job / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a, a.ecs_per_gpu), "n_groups": N_GROUPS, "group_size": GROUP_SIZE,
                   "ecs_per_gpu": a.ecs_per_gpu, "algo": a.algo},
        "cpu_baseline": {**cb, "value": value},
        "e2e": {"value": value, "unit": "EC-iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.time() - t0,
    }
    print(json.dumps(line), flush=True)


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
        sess = lik.vi_begin(algo, tol=0.0 if a.algo == "em" else -1e300, max_iters=10 ** 9, time_kernels=True)
        sess.step(a.warmup)
        st0 = sess.poll()
        torch.cuda.synchronize()
        dist.barrier()
        sampler = ClockSampler(local) if rank == 0 else None
        launches0 = M.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        sess.step(a.steps)
        ev1.record(stream)
        torch.cuda.synchronize()
        dist.barrier()
        launches = M.launch_count() - launches0
        clocks = sampler.stop() if sampler else None
        ms_total = dist.reduce_max(ev0.elapsed_time(ev1))
        st1 = sess.poll()
        res = sess.finish()
    assert st1.iters - st0.iters == a.steps, "the timed region must contain exactly --steps iterations"
    pass_ms = (st1.pass_ms_sum - st0.pass_ms_sum) / max(1, st1.pass_launches - st0.pass_launches)
    passes_per_iter = (st1.pass_launches - st0.pass_launches) / a.steps
    bytes_per_launch = st1.pass_bytes / passes_per_iter          # pass_bytes is per iteration (all sweeps of it)
    n_job = dist.reduce_sum(n_ecs_local)
    ms_per_step = ms_total / a.steps
    value = n_job * a.steps / (ms_total * 1e-3)
    theta_sum = float(res.theta.sum())
    lik.close(); aln.close()

    # ---- end-to-end leg: host buffers -> abundances through the C ABI, copies inside the timed region --
    e2e = None
    if not a.no_e2e:
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        with torch.cuda.stream(stream):
            aln = M.Alignment(ctx, wl.n_reads, wl.n_targets, wl.row_ptr, wl.targets, partitioned=world > 1)
            lik = M.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes, storage=storage)
            r2 = lik.vi_run(algo, tol=0.0 if a.algo == "em" else -1e300, max_iters=a.steps, poll_every=a.steps)
        torch.cuda.synchronize()
        sec = dist.reduce_max(time.perf_counter() - t0)
        assert r2.iters == a.steps
        h2d = wl.row_ptr.nbytes + wl.targets.nbytes + wl.group_of_target.nbytes + wl.group_sizes.nbytes + 8 * N_GROUPS
        e2e = {"value": n_job * a.steps / sec, "unit": "EC-iter/s", "vi_iters_per_s": a.steps / sec, "seconds": sec,
               "h2d_bytes_per_step": h2d / a.steps, "d2h_bytes_per_step": (8 * N_GROUPS + 64) / a.steps,
               "what": "mswb_ec_build + mswb_lik_build + mswb_vi_run(K iterations) from host CSR buffers, theta back on the host",
               "allocator": "the library's device block cache is warm (the resident leg ran first): a cold process pays the "
                            "cudaMalloc of the matrix once on top (0.2-0.4 s at 100 GB)"}
        lik.close(); aln.close()

    # ---- side series on the same inputs (N = 1 only): the fp64 forms of the sweep, which cannot hold the full shard ----
    extras = None
    if world == 1 and not a.no_extras and a.storage == "f32" and a.algo == "em":
        extras = {}
        peak, _ = peak_gbs()
        for name, st, al, n_sub, steps in (("em_f64", M.STORE_F64, M.ALGO_EM, min(n_local, 6_000_000), 20),
                                           ("rcg_f64", M.STORE_F64, M.ALGO_RCG, min(n_local, 1_000_000), 10)):
            rp = wl.row_ptr[:n_sub + 1]
            with torch.cuda.stream(stream):
                aln = M.Alignment(ctx, n_sub, wl.n_targets, rp, wl.targets[:int(rp[-1])])
                lik = M.Likelihood.build(ctx, aln, wl.group_of_target, wl.group_sizes, storage=st)
                sess = lik.vi_begin(al, tol=0.0 if al == M.ALGO_EM else -1e300, max_iters=10 ** 9, time_kernels=True)
                sess.step(a.warmup)
                s0 = sess.poll()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                sess.step(steps)
                e1.record(stream)
                torch.cuda.synchronize()
                s1 = sess.poll()
                sess.finish()
            ms = e0.elapsed_time(e1) / steps
            kms = (s1.pass_ms_sum - s0.pass_ms_sum) / steps
            gbs = s1.pass_bytes / (kms * 1e-3) / 1e9
            extras[name] = {"ecs": lik.n_ecs, "n_groups": N_GROUPS, "steps": steps, "ms_per_step": ms, "vi_iters_per_s": 1e3 / ms,
                            "pass_kernels_ms_per_step": kms, "achieved_gbs": gbs, "frac_of_measured_peak": gbs / peak,
                            "frac_of_nominal_8TBs": gbs / 8000.0, "bytes_per_step": s1.pass_bytes}
            lik.close(); aln.close()

    cb = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cb = cpu_baseline(wl, a, n_local, steps=3, warmup=1)

    if rank == 0:
        peak, peak_src = peak_gbs()
        achieved = bytes_per_launch / (pass_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                tj = json.load(f).get(f"{a.algo}_{a.storage}")
            if tj:   # dram bytes per element from the ncu --set full capture, scaled to this launch
                traffic = tj["dram_bytes_per_element"] * n_ecs_local * N_GROUPS
        line = {
            "metric": "VI throughput (EC-iterations/s)", "value": value, "unit": "EC-iter/s",
            "vi_iters_per_s": a.steps / (ms_total * 1e-3), "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 accumulation / f32 storage" if a.storage == "f32" else "f64", "data": "synthetic",
            "config": {"workload": workload_name(a, n_local), "n_groups": N_GROUPS, "group_size": GROUP_SIZE,
                       "ecs_per_gpu": n_ecs_local, "ecs_job": n_job, "algo": a.algo, "storage": a.storage,
                       "l2": "inputs larger than L2 (shard >> 126 MB), no flush needed", "parallelism": f"ec-shard x{world}",
                       "generator_s": round(t_gen, 1)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "frac_of_nominal_8TBs": achieved / 8000.0, "peak_source": peak_src, "traffic": traffic,
                         "kernel": ("em_sparse_pass_kernel" if a.storage == "sparse" else "em_lin_pass_kernel") if a.algo == "em" else "rcg_sweep_a_kernel + rcg_sweep_b_kernel",
                         "bytes_per_launch": bytes_per_launch, "kernel_ms": pass_ms,
                         "kernel_share_of_step": pass_ms * passes_per_iter / ms_per_step},
            "cpu_baseline": cb,
            "e2e": e2e,
            "gpu_launches": launches,
            "clocks": clocks,
            "extras": extras,
            "check": {"theta_sum": theta_sum, "bound": res.bound},
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    dist.finalize()


if __name__ == "__main__":
    main()
