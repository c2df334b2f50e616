#!/usr/bin/env python
"""Static model of the L1/shared-memory data-pipe work per tile of the two tile-kernel variants.

ncu shows the LSU data pipe (l1tex__data_pipe_lsu_wavefronts) as the busiest unit of the tile
kernel (81-90 %, profiles/ncu_r01_*.txt).  This script counts, for real tiles of the BASELINE
workloads, the wavefronts each variant sends through that pipe:

  * shared-memory loads/stores: one wavefront per conflict-free 128-byte phase; a warp access costs
    max over the 32 banks of the number of DISTINCT 4-byte words requested in that bank (64-bit
    accesses are issued as two half-warps);
  * x gathers: one wavefront per distinct 128-byte line touched by a warp's LDG (the t-stage cost;
    sectors are counted too).

variant 2 (shipped, spmv_tile.cuh): column indices / values / products accessed strip-mined
(conflict-free), products re-read thread-blocked in the walk, popcount prefix reads the bitmap.
variant 3 (spmv_tile3.cuh): column indices and values read thread-blocked once, no product
round trip, start rows scattered by the row owners.

It runs on the CPU from the generators alone (no GPU): python tools/lsu_model.py [--tiles N]
The numbers are a model, not a measurement; they rank access patterns, they do not predict time.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

THREADS = 128


def wavefronts_32(addr_words, active):
    """addr_words: int array [32] of 4-byte word addresses; active: bool [32]."""
    if not active.any():
        return 0
    w = np.unique(addr_words[active])
    return int(np.bincount(w % 32, minlength=32).max())


def wavefronts_64(addr_elems, active):
    """8-byte elements: two half-warps, each element covers two consecutive banks."""
    total = 0
    for h in (slice(0, 16), slice(16, 32)):
        a, m = addr_elems[h], active[h]
        if not m.any():
            continue
        e = np.unique(a[m])
        total += int(np.bincount(e % 16, minlength=16).max())
    return total


def lines_and_sectors(cols, active, vbytes):
    if not active.any():
        return 0, 0
    b = cols[active].astype(np.int64) * vbytes
    return len(np.unique(b // 128)), len(np.unique(b // 32))


def model_tile(flags, cols, ipt, vbytes):
    """flags: uint8 [TILE] (1 = row end) in merge order; cols: column index of every nonzero of the
    tile in order.  Returns dict variant -> dict of wavefront counts."""
    tile = THREADS * ipt
    nnz = int((flags == 0).sum())
    nz_index = np.cumsum(flags == 0) - 1            # nonzero index of every merge item (valid where flag == 0)
    ends_before = np.concatenate([[0], np.cumsum(flags)[:-1]])
    wf = wavefronts_64 if vbytes == 8 else wavefronts_32
    out = {2: dict(smem=0, gather_lines=0, gather_sectors=0), 3: dict(smem=0, gather_lines=0, gather_sectors=0)}
    lanes = np.arange(32)
    for warp in range(THREADS // 32):
        tids = warp * 32 + lanes
        # ---------------- variant 2 ----------------
        v2 = 0
        for i in range(ipt):                       # strip-mined: j = tid + i*128
            j = tids + i * THREADS
            act = j < nnz
            v2 += wavefronts_32(j, act)            # s_col load
            v2 += 2 * wf(j, act)                   # s_val load + product store
            l, s = lines_and_sectors(cols[np.minimum(j, nnz - 1)] if nnz else j, act, vbytes)
            out[2]["gather_lines"] += l
            out[2]["gather_sectors"] += s
        for i in range(ipt):                       # walk: thread-blocked product reads
            p = tids * ipt + i
            act = (p < tile) & (flags[np.minimum(p, tile - 1)] == 0)
            v2 += wf(nz_index[np.minimum(p, tile - 1)], act)
        v2 += 2 + 2 + (ipt - 1)                    # bitmap words: own two, before_warp strip, in_warp broadcast reads
        out[2]["smem"] += v2
        # ---------------- variant 3 ----------------
        v3 = 0
        for i in range(ipt):                       # thread-blocked column index + value reads
            p = tids * ipt + i
            act = (p < tile) & (flags[np.minimum(p, tile - 1)] == 0)
            k = nz_index[np.minimum(p, tile - 1)]
            v3 += wavefronts_32(k, act) + wf(k, act)
            l, s = lines_and_sectors(cols[np.clip(k, 0, max(nnz - 1, 0))] if nnz else k, act, vbytes)
            out[3]["gather_lines"] += l
            out[3]["gather_sectors"] += s
        v3 += 2 + 1                                # bitmap words, s_xs
        out[3]["smem"] += v3
    # row-end loop (both): s_row loads + atomicOr; variant 3 also the previous offset and the s_xs scatter
    nrows = int(flags.sum())
    pos = np.nonzero(flags)[0]
    for base in range(0, nrows, THREADS):
        for warp in range(THREADS // 32):
            r = base + warp * 32 + lanes
            act = r < nrows
            if not act.any():
                continue
            rr = np.minimum(r, nrows - 1)
            a = wavefronts_32(rr, act)
            words = pos[rr] // 32
            atom = int(np.bincount(words[act] % 32, minlength=32).max())  # same-word atomics serialise too
            out[2]["smem"] += a + atom
            out[3]["smem"] += 2 * a + atom
            prev = np.where(rr > 0, pos[np.maximum(rr - 1, 0)], -1)
            span = pos[rr] // ipt - (prev + ipt) // ipt + 1
            out[3]["smem"] += int(np.maximum(span[act], 0).max())       # serial s_xs stores of the longest range in the warp
    return out


def sample_tiles(name, n_tiles, rng):
    import torch
    from merge_spmv_b200 import generators as gen

    kind, dt, p = gen.CONFIGS[name]
    vbytes = 8 if dt == torch.float64 else 4
    ipt = 9 if vbytes == 8 else 13
    tile = THREADS * ipt
    if kind in ("uniform", "uniform_local"):
        ro = gen.uniform_row_offsets(p["rows"], p["nnz_per_row"])
        cols_n = p["cols"]
        gkind = "local" if kind == "uniform_local" else "stratified"
    elif kind == "powerlaw":
        lengths, _ = gen.powerlaw_row_lengths(p["rows"], min(p["max_row"], p["cols"]), p["target_nnz"])
        ro = gen._offsets_from_lengths(lengths)
        cols_n = p["cols"]
        gkind = "stratified"
    else:
        ro = gen.banded_row_offsets(p["rows"], p["half_bandwidth"])
        cols_n = p["rows"]
        gkind = "banded"
    ro_np = ro.numpy().astype(np.int64)
    rows, nnz = ro_np.size - 1, int(ro_np[-1])
    total = rows + nnz
    n_all = (total + tile - 1) // tile
    picks = rng.choice(n_all - 1, size=min(n_tiles, n_all - 1), replace=False)
    row_end_diag = ro_np[1:] + np.arange(rows)      # merge position of every row end
    seed = {"uniform": 0x5EED0001, "uniform_local": 0x5EED0001, "powerlaw": 0x5EED0003, "banded": 0x5EED0004}[kind]
    for t in picks:
        d0, d1 = t * tile, (t + 1) * tile
        r0 = int(np.searchsorted(row_end_diag, d0, side="left"))
        r1 = int(np.searchsorted(row_end_diag, d1, side="left"))
        flags = np.zeros(tile, np.uint8)
        flags[row_end_diag[r0:r1] - d0] = 1
        y0 = d0 - r0
        k1 = y0 + int((flags == 0).sum())
        col, _ = gen.fill_nonzeros(ro, cols_n, y0, k1, kind=gkind, dtype=dt, values="ones", device="cpu",
                                   half_bandwidth=p.get("half_bandwidth", 3), seed=seed)
        yield flags, col.numpy(), ipt, vbytes


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tiles", type=int, default=24)
    ap.add_argument("--workloads", default="uniform_1m_64,powerlaw_2m,banded_10m,uniform_1m_64_local")
    args = ap.parse_args()
    rng = np.random.default_rng(7)
    print(f"{'workload':22s} {'variant':>7s} {'smem wf/tile':>13s} {'gather lines':>13s} {'gather sectors':>15s} {'total wf':>9s}")
    for name in args.workloads.split(","):
        acc = {2: np.zeros(3), 3: np.zeros(3)}
        n = 0
        for flags, cols, ipt, vbytes in sample_tiles(name, args.tiles, rng):
            m = model_tile(flags, cols, ipt, vbytes)
            for v in (2, 3):
                acc[v] += [m[v]["smem"], m[v]["gather_lines"], m[v]["gather_sectors"]]
            n += 1
        for v in (2, 3):
            s, l, sec = acc[v] / n
            print(f"{name:22s} {v:7d} {s:13.0f} {l:13.0f} {sec:15.0f} {s + l:9.0f}")


if __name__ == "__main__":
    main()
