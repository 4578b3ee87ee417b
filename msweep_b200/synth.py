"""Seeded synthetic Themisto-style pseudoalignments (SURVEY.md §8(d) "configs restated").

The reference ships no data (its toy set is a Zenodo download), so every workload here is
generated: a grouping of T reference sequences into K lineages, a truth abundance vector over a
few present lineages, and reads whose hit pattern follows the LL_WOR21 model's own story — the
source lineage's sequences are hit with probability q, a few related lineages with a small
probability, everything else (almost) never.  Reads are drawn from a pool of pattern templates
so that the number of distinct patterns (equivalence classes) is controllable.

Pure numpy; used by tests/, bench.py and the examples.  No dependency on the oracle.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np


@dataclass
class Workload:
    n_reads: int
    n_targets: int
    n_groups: int
    row_ptr: np.ndarray          # uint64 [R+1]  CSR over reads (already strand-merged)
    targets: np.ndarray          # uint32 [nnz]  ascending inside each read
    group_of_target: np.ndarray  # uint32 [T]
    group_sizes: np.ndarray      # uint64 [K]
    group_names: list[str]
    truth: np.ndarray            # float64 [K] generating abundances


def make_grouping(n_targets: int, n_groups: int, layout: str, rng: np.random.Generator):
    """group ids are assigned in order of first appearance, as the reference does
    (include/Grouping.hpp:62-67), so `group_of_target` is exactly what reading the -i file yields."""
    if layout == "block":
        label = np.arange(n_targets) * n_groups // n_targets
    elif layout == "interleaved":
        label = np.arange(n_targets) % n_groups
    elif layout == "shuffled":
        label = rng.permutation(np.arange(n_targets) % n_groups)
    else:
        raise ValueError(layout)
    _, first = np.unique(label, return_index=True)
    order = np.argsort(first)                 # labels in order of first appearance
    remap = np.empty(len(order), np.int64)
    remap[order] = np.arange(len(order))
    got = remap[label].astype(np.uint32)
    names = [f"g{int(l):05d}" for l in np.arange(len(order))[order]]
    sizes = np.bincount(got, minlength=len(order)).astype(np.uint64)
    return got, sizes, names


def generate(n_reads: int, n_targets: int, n_groups: int, *, n_present: int = 5, n_templates: int = 2000,
             n_related: int = 3, q_hit: float = 0.65, q_related: float = 0.15, p_noise: float = 0.02,
             p_unaligned: float = 0.05, layout: str = "shuffled", seed: int = 20231017,
             chunk: int = 1 << 18) -> Workload:
    rng = np.random.Generator(np.random.Philox(seed))
    got, sizes, names = make_grouping(n_targets, n_groups, layout, rng)
    K = len(sizes)
    # members[g] = ascending target ids of group g, padded to the max size with -1
    S = int(sizes.max())
    order = np.argsort(got, kind="stable")
    starts = np.concatenate(([0], np.cumsum(sizes)[:-1])).astype(np.int64)
    members = np.full((K, S), -1, np.int64)
    slot = np.arange(n_targets) - np.repeat(starts, sizes.astype(np.int64))
    members[got[order], slot] = order

    present = rng.choice(K, size=min(n_present, K), replace=False)
    theta = rng.dirichlet(np.ones(len(present)))
    truth = np.zeros(K)
    truth[present] = theta
    related = np.stack([rng.choice(K, size=min(n_related, K - 1) + 1, replace=False) for _ in range(K)])

    # ---- template pool: n_templates patterns per present lineage --------------------------------
    n_pool = len(present) * n_templates
    src = np.repeat(present, n_templates)
    L = 1 + min(n_related, K - 1)
    t_ptr = [np.zeros(1, np.int64)]
    t_tgt = []
    base = 0
    thr = np.empty((L,), np.uint8)
    thr[0] = int(q_hit * 256)
    thr[1:] = int(q_related * 256)
    for c0 in range(0, n_pool, chunk):
        s = src[c0:c0 + chunk]
        lin = related[s][:, :L].copy()
        clash = lin[:, 1:] == s[:, None]
        lin[:, 0], lin[:, 1:] = s, np.where(clash, related[s][:, :1], lin[:, 1:])
        cand = members[lin]                                            # [n, L, S] target ids or -1
        hit = rng.integers(0, 256, size=cand.shape, dtype=np.uint8) < thr[None, :, None]
        hit &= cand >= 0
        # duplicates (a related lineage drawn twice) are removed by the per-row sort/unique below
        rows, li, si = np.nonzero(hit)
        tg = cand[rows, li, si]
        key = np.lexsort((tg, rows))
        rows, tg = rows[key], tg[key]
        keep = np.ones(len(tg), bool)
        keep[1:] = (rows[1:] != rows[:-1]) | (tg[1:] != tg[:-1])
        rows, tg = rows[keep], tg[keep]
        lens = np.bincount(rows, minlength=len(s))
        t_ptr.append(base + np.cumsum(lens))
        base += int(lens.sum())
        t_tgt.append(tg.astype(np.uint32))
    t_ptr = np.concatenate(t_ptr)
    t_tgt = np.concatenate(t_tgt) if t_tgt else np.zeros(0, np.uint32)

    # ---- reads: pick a source lineage by theta, then one of its templates -------------------------
    which = rng.choice(len(present), size=n_reads, p=theta)
    tmpl = which * n_templates + rng.integers(0, n_templates, size=n_reads)
    unaligned = rng.random(n_reads) < p_unaligned
    noisy = (rng.random(n_reads) < p_noise) & ~unaligned
    noise_t = rng.integers(0, n_targets, size=n_reads).astype(np.uint32)
    lens = (t_ptr[tmpl + 1] - t_ptr[tmpl]).astype(np.int64)
    lens[unaligned] = 0
    # noisy reads (a few percent; rows are short): template row plus one extra target, kept ascending/unique
    noisy_idx = np.nonzero(noisy)[0]
    noisy_rows = [np.unique(np.append(t_tgt[t_ptr[tmpl[r]]:t_ptr[tmpl[r] + 1]], noise_t[r])) for r in noisy_idx]
    out_len = lens.copy()
    out_len[noisy_idx] = [len(x) for x in noisy_rows]
    row_ptr = np.zeros(n_reads + 1, np.uint64)
    row_ptr[1:] = np.cumsum(out_len)
    targets = np.empty(int(row_ptr[-1]), np.uint32)
    rp = row_ptr[:-1].astype(np.int64)
    glens = lens.copy()
    glens[noisy_idx] = 0
    rep = np.repeat(np.arange(n_reads), glens)
    within = np.arange(int(glens.sum())) - np.repeat(np.cumsum(glens) - glens, glens)
    targets[rp[rep] + within] = t_tgt[t_ptr[tmpl][rep] + within]
    for r, row in zip(noisy_idx, noisy_rows):
        targets[rp[r]:rp[r] + len(row)] = row
    return Workload(n_reads, n_targets, K, row_ptr, targets, got, sizes, names, truth)


def generate_ec_patterns(n_patterns: int, n_groups: int, group_size: int, *, n_present: int = 50, n_related: int = 3,
                         q_hit: float = 0.65, q_related: float = 0.15, seed: int = 20231019,
                         chunk: int = 1 << 20, dup_factor: float = 0.0, n_pool: int = 0, workers: int | None = None) -> Workload:
    """Bench-scale generator (config 3): every read gets its own freshly drawn pattern, block grouping
    (targets of lineage g are g*S .. g*S+S-1) so rows come out sorted without a sort.  With
    dup_factor > 0 a fraction of reads repeat the previous read's pattern (EC counts > 1).  With n_pool > 0 the
    related lineages come from a fixed pool of that many lineages (the present ones included), so that every other
    lineage receives no hit at all (config 4: most lineages empty, --min-hits 1 prunes them)."""
    rng = np.random.Generator(np.random.Philox(seed))
    K, S = n_groups, group_size
    T = K * S
    got = (np.arange(T) // S).astype(np.uint32)
    sizes = np.full(K, S, np.uint64)
    names = [f"g{g:05d}" for g in range(K)]
    present = np.sort(rng.choice(K, size=min(n_present, K), replace=False))
    theta = rng.dirichlet(np.ones(len(present)))
    truth = np.zeros(K)
    truth[present] = theta
    L = 1 + n_related
    pool = None
    if n_pool > 0:
        others = np.setdiff1d(np.arange(K), present)
        extra = rng.choice(others, size=max(0, min(n_pool, K) - len(present)), replace=False)
        pool = np.sort(np.concatenate([present, extra]))
    thr_src, thr_rel = int(q_hit * 256), int(q_related * 256)
    n_chunks = (n_patterns + chunk - 1) // chunk

    def one_chunk(ci: int):
        # every chunk draws from its own Philox stream (jumped 2^128 draws apart): chunks can be generated in any
        # order, on any number of threads, with the same result
        crng = np.random.Generator(np.random.Philox(seed).jumped(ci + 1))
        n = min(chunk, n_patterns - ci * chunk)
        s = present[crng.choice(len(present), size=n, p=theta)]
        if pool is None:
            rel = (s[:, None] + crng.integers(1, K, size=(n, n_related))) % K
        else:
            rel = pool[crng.integers(0, len(pool), size=(n, n_related))]
        lin = np.sort(np.concatenate([s[:, None], rel], axis=1), axis=1)              # ascending lineages
        is_src = lin == s[:, None]
        thr = np.where(is_src, thr_src, thr_rel).astype(np.uint8)
        hit = crng.integers(0, 256, size=(n, L, S), dtype=np.uint8) < thr[:, :, None]
        dup = np.zeros((n, L), bool)
        dup[:, 1:] = lin[:, 1:] == lin[:, :-1]
        hit &= ~dup[:, :, None]
        rows, li, si = np.nonzero(hit)                                                 # row-major => sorted
        tg = (lin[rows, li] * S + si).astype(np.uint32)
        return np.bincount(rows, minlength=n).astype(np.uint64), tg

    if workers is None:
        workers = min(n_chunks, os.cpu_count() or 1)
    if workers > 1 and n_chunks > 1:
        from concurrent.futures import ThreadPoolExecutor        # numpy releases the GIL in the heavy calls
        with ThreadPoolExecutor(max_workers=workers) as ex:
            parts = list(ex.map(one_chunk, range(n_chunks)))
    else:
        parts = [one_chunk(ci) for ci in range(n_chunks)]
    row_ptr = np.zeros(n_patterns + 1, np.uint64)
    np.cumsum(np.concatenate([p[0] for p in parts]), out=row_ptr[1:])
    targets = np.concatenate([p[1] for p in parts])
    del parts
    wl = Workload(n_patterns, T, K, row_ptr, targets, got, sizes, names, truth)
    if dup_factor > 0:
        wl = _duplicate_reads(wl, dup_factor, rng)
    return wl


def _duplicate_reads(wl: Workload, dup_factor: float, rng) -> Workload:
    reps = 1 + rng.geometric(1.0 / (1.0 + dup_factor), size=wl.n_reads) - 1
    lens = (wl.row_ptr[1:] - wl.row_ptr[:-1]).astype(np.int64)
    new_lens = np.repeat(lens, reps)
    row_ptr = np.zeros(len(new_lens) + 1, np.uint64)
    row_ptr[1:] = np.cumsum(new_lens)
    src_start = np.repeat(wl.row_ptr[:-1].astype(np.int64), reps)
    within = np.arange(int(new_lens.sum())) - np.repeat(row_ptr[:-1].astype(np.int64), new_lens)
    targets = wl.targets[np.repeat(src_start, new_lens) + within]
    return Workload(len(new_lens), wl.n_targets, wl.n_groups, row_ptr, targets, wl.group_of_target,
                    wl.group_sizes, wl.group_names, wl.truth)


def write_grouping(path: str, wl: Workload) -> None:
    with open(path, "w") as f:
        for g in wl.group_of_target:
            f.write(wl.group_names[int(g)] + "\n")


def _format_tokens(vals: np.ndarray, sep: np.ndarray) -> np.ndarray:
    """ASCII bytes of `vals` in decimal, token i followed by the byte sep[i] (vectorised; no Python loop per token)."""
    v = vals.astype(np.int64)
    nd = np.ones(len(v), np.int64)
    p10 = 10
    while True:
        m = v >= p10
        if not m.any():
            break
        nd += m
        p10 *= 10
    end = np.cumsum(nd + 1)
    buf = np.empty(int(end[-1]) if len(end) else 0, np.uint8)
    buf[end - 1] = sep
    d = 0
    while True:
        m = nd > d
        if not m.any():
            break
        buf[end[m] - 2 - d] = 48 + v[m] % 10
        v //= 10
        d += 1
    return buf


def write_themisto(prefix: str, wl: Workload, *, paired: bool = True, seed: int = 7, shuffle_frac: float = 0.01,
                   rows_per_block: int = 1 << 16):
    """Writes Themisto plaintext ("<read_id> <t0> <t1> ...").  Paired: strand k = merged pattern plus
    strand-private extra hits (odd targets only on strand 1, even only on strand 2, inserted at a random place:
    Themisto does not sort the targets inside a line), so that the default intersection merge recovers `wl` exactly.
    A small fraction of lines is emitted out of order (Themisto's multi-threaded output is unordered;
    include/mSWEEP_alignment.hpp:60-64 indexes by the id column, not the line number).  Vectorised: 1e6 reads x 2
    strands (600 MB of text) take seconds, not a minute."""
    rng = np.random.Generator(np.random.Philox(seed))
    R, T = wl.n_reads, wl.n_targets
    paths = [f"{prefix}_1.aln", f"{prefix}_2.aln"] if paired else [f"{prefix}.aln"]
    rp = wl.row_ptr.astype(np.int64)
    lens = np.diff(rp)
    keys = np.repeat(np.arange(R, dtype=np.int64), lens) * T + wl.targets.astype(np.int64)     # ascending (rows are sorted)
    for k, p in enumerate(paths):
        order = np.arange(R)
        n_sw = int(R * shuffle_frac) // 2 * 2
        if n_sw:
            pick = rng.choice(R, size=n_sw, replace=False)
            order[pick] = order[pick[::-1]]
        # strand-private extra hit for ~30 % of the aligned reads
        extra = np.full(R, -1, np.int64)
        if paired and T >= 2:
            cand = (lens > 0) & (rng.random(R) < 0.3)
            e = rng.integers(0, T // 2, size=R) * 2 + (1 - k)
            ok = cand & (e < T)
            idx = np.flatnonzero(ok)
            pos = np.searchsorted(keys, idx * T + e[idx])
            present = (pos < len(keys)) & (keys[np.minimum(pos, len(keys) - 1)] == idx * T + e[idx])
            extra[idx[~present]] = e[idx[~present]]
        ins_pos = (rng.random(R) * (lens + 1)).astype(np.int64)                                 # where the extra hit goes
        def block(b0: int) -> bytes:
            rows = order[b0:b0 + rows_per_block]
            n = len(rows)
            has = extra[rows] >= 0
            ln = lens[rows] + has                       # targets on the line
            tok_per = ln + 1                            # + the read id
            tend = np.cumsum(tok_per)
            tstart = tend - tok_per
            vals = np.empty(int(tend[-1]), np.int64)
            sep = np.full(len(vals), 32, np.uint8)
            sep[tend - 1] = 10
            vals[tstart] = rows
            # original targets: slot = their index in the row, shifted by one behind the insertion point
            rep = np.repeat(np.arange(n), lens[rows])
            within = np.arange(int(lens[rows].sum())) - np.repeat(np.cumsum(lens[rows]) - lens[rows], lens[rows])
            shift = (has[rep] & (within >= ins_pos[rows][rep])).astype(np.int64)
            vals[tstart[rep] + 1 + within + shift] = wl.targets[rp[rows][rep] + within]
            hi = np.flatnonzero(has)
            vals[tstart[hi] + 1 + ins_pos[rows][hi]] = extra[rows][hi]
            return _format_tokens(vals, sep).tobytes()

        from concurrent.futures import ThreadPoolExecutor        # numpy releases the GIL in the heavy calls
        with open(p, "wb") as f, ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
            for chunk_bytes in ex.map(block, range(0, R, rows_per_block)):
                f.write(chunk_bytes)
    return paths
