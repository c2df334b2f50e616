"""Synthetic CSR inputs for the BASELINE.json configs, plus the reference's own generators.

The reference only has dense / grid2d / grid3d / wheel (sparse_matrix.h:386-617); uniform,
power-law and banded matrices are builder-defined.  Everything here is
counter-based (splitmix64 of seed + index), written with wrapping int64 torch ops, so the same
matrix comes out bit-identical on CPU and on any CUDA device, for any sub-range of nonzeros
(a rank can generate just its own shard).

Column model for uniform / power-law rows: *stratified sampling without replacement* -- the
j-th nonzero of a row of length L over C columns is drawn uniformly from stratum
[floor(j*C/L), floor((j+1)*C/L)).  Columns are therefore sorted, distinct and marginally uniform
over [0, C); for short rows (L << C) consecutive nonzeros are ~C/L apart, i.e. as cache-hostile
as i.i.d. uniform columns.  Values: "ones" (the reference drivers' choice, gpu_spmv.cu:521-525 --
y is then exactly the row length in any summation order) or "random" = U[0.5, 1.5).
"""
from __future__ import annotations

import math

import numpy as np
import torch

_M64 = (1 << 64) - 1


def _s64(u):
    u &= _M64
    return u - (1 << 64) if u >= (1 << 63) else u


_C1, _C2, _C3 = _s64(0x9E3779B97F4A7C15), _s64(0xBF58476D1CE4E5B9), _s64(0x94D049BB133111EB)


def _lsr(z, s):
    return (z >> s) & ((1 << (64 - s)) - 1)


def splitmix64(z: torch.Tensor) -> torch.Tensor:
    """splitmix64 finaliser on int64 tensors (two's-complement wraparound == uint64 math)."""
    z = z + _C1
    z = (z ^ _lsr(z, 30)) * _C2
    z = (z ^ _lsr(z, 27)) * _C3
    return z ^ _lsr(z, 31)


def _uniform_half_to_three_halves(index: torch.Tensor, seed: int, dtype) -> torch.Tensor:
    h = splitmix64(index + _s64(seed * 0x100000001B3))
    return (0.5 + _lsr(h, 11).to(torch.float64) * (1.0 / 9007199254740992.0)).to(dtype)


def vector(n, dtype=torch.float64, mode="ones", seed=0x5EED00FF, device="cpu"):
    """x vector: all ones (gpu_spmv.cu:521-522) or U[0.5,1.5) hashed per index."""
    if mode == "ones":
        return torch.ones(n, dtype=dtype, device=device)
    idx = torch.arange(n, dtype=torch.int64, device=device)
    return _uniform_half_to_three_halves(idx, seed, dtype)


# ---------------------------------------------------------------------------------------------
# row structures
# ---------------------------------------------------------------------------------------------
def _offsets_from_lengths(lengths: torch.Tensor) -> torch.Tensor:
    ro = torch.zeros(lengths.numel() + 1, dtype=torch.int64)
    torch.cumsum(lengths.to(torch.int64), 0, out=ro[1:])
    if int(ro[-1]) + lengths.numel() >= 2 ** 31 - 65536:
        raise ValueError("rows + nnz must stay below 2^31 (int32 offsets, dispatch_spmv_orig.cuh:608)")
    return ro.to(torch.int32)


def uniform_row_offsets(rows, nnz_per_row):
    return (torch.arange(rows + 1, dtype=torch.int64) * nnz_per_row).to(torch.int32)


# bisection results for the two BASELINE.json power-law configs (saves 60 passes over 2-20 M rows)
_KNOWN_ALPHA = {
    (2_000_000, 1_000_000, 200_000_000): 0.7215631890166361,
    (20_000_000, 1_000_000, 1_000_000_000): 0.6509830663579996,
}


def powerlaw_row_lengths(rows, max_row, target_nnz, seed=0x5EED0003):
    """len(k) = max(1, floor(max_row / k^alpha)), k = 1..rows, alpha solved so the sum hits
    target_nnz; rank -> row by a seeded permutation.  Returns (lengths int64[rows], alpha)."""
    k = torch.arange(1, rows + 1, dtype=torch.float64)

    def total(alpha):
        return int(torch.clamp(torch.floor(max_row / torch.pow(k, alpha)), min=1).sum().item())

    alpha = _KNOWN_ALPHA.get((rows, max_row, target_nnz))
    if alpha is None:
        lo, hi = 0.0, 4.0
        for _ in range(60):
            mid = 0.5 * (lo + hi)
            if total(mid) > target_nnz:
                lo = mid
            else:
                hi = mid
        alpha = hi
    lengths = torch.clamp(torch.floor(max_row / torch.pow(k, alpha)), min=1).to(torch.int64)
    perm = torch.argsort(splitmix64(torch.arange(rows, dtype=torch.int64) + _s64(seed)))
    out = torch.empty_like(lengths)
    out[perm] = lengths  # rank r lands on row perm[r]
    return out, alpha


def banded_row_offsets(rows, half_bandwidth):
    r = torch.arange(rows, dtype=torch.int64)
    lengths = torch.clamp(r + half_bandwidth, max=rows - 1) - torch.clamp(r - half_bandwidth, min=0) + 1
    return _offsets_from_lengths(lengths)


# ---------------------------------------------------------------------------------------------
# nonzeros of an index range [k0, k1) -- chunked so that 10^9-nonzero shards fit comfortably
# ---------------------------------------------------------------------------------------------
def _rows_of(row_offsets_dev64, k):
    return torch.searchsorted(row_offsets_dev64, k, right=True) - 1


def fill_nonzeros(row_offsets, cols, k0, k1, *, kind, dtype, values="ones", seed=0x5EED0001,
                  half_bandwidth=3, device="cpu", chunk=1 << 25, out_col=None, out_val=None, window=4096):
    """Column indices (int32) and values of nonzeros k0..k1-1 of the matrix whose full
    row_offsets is given.  kind: "stratified" (uniform & power-law rows: columns over all of
    [0, cols)), "local" (same stratified draw, but inside a `window`-wide column window centred on
    the row -- FEM-like locality) or "banded"."""
    n = k1 - k0
    col = out_col if out_col is not None else torch.empty(n, dtype=torch.int32, device=device)
    val = out_val if out_val is not None else torch.empty(n, dtype=dtype, device=device)
    ro64 = row_offsets.to(device=device, dtype=torch.int64)
    for c0 in range(k0, k1, chunk):
        c1 = min(c0 + chunk, k1)
        k = torch.arange(c0, c1, dtype=torch.int64, device=device)
        row = _rows_of(ro64, k)
        start = ro64[row]
        j = k - start
        if kind == "banded":
            c = torch.clamp(row - half_bandwidth, min=0) + j
        elif kind == "stratified":
            length = ro64[row + 1] - start
            lo = (j * cols) // length
            hi = ((j + 1) * cols) // length
            h = _lsr(splitmix64(k + _s64(seed)), 1)
            c = lo + h % torch.clamp(hi - lo, min=1)
        elif kind == "local":
            length = ro64[row + 1] - start
            w = min(window, cols)
            w0 = torch.clamp(row * cols // max(ro64.numel() - 1, 1) - w // 2, min=0, max=cols - w)
            lo = (j * w) // length
            hi = ((j + 1) * w) // length
            h = _lsr(splitmix64(k + _s64(seed)), 1)
            c = w0 + lo + h % torch.clamp(hi - lo, min=1)
        else:
            raise ValueError(kind)
        col[c0 - k0:c1 - k0] = c.to(torch.int32)
        if values == "ones":
            val[c0 - k0:c1 - k0] = 1
        else:
            val[c0 - k0:c1 - k0] = _uniform_half_to_three_halves(k, seed ^ 0xABCDEF, dtype)
        del k, row, start, j, c
    return col, val


class Csr:
    """A CSR matrix as three tensors + shape (row_offsets int32[rows+1], col int32, val)."""

    def __init__(self, rows, cols, row_offsets, col, val, name=""):
        self.rows, self.cols = int(rows), int(cols)
        self.row_offsets, self.col, self.val = row_offsets, col, val
        self.name = name

    @property
    def nnz(self):
        return int(self.val.numel())

    def to(self, device):
        return Csr(self.rows, self.cols, self.row_offsets.to(device), self.col.to(device),
                   self.val.to(device), self.name)

    def numpy(self):
        return (self.row_offsets.cpu().numpy(), self.col.cpu().numpy(), self.val.cpu().numpy())

    def algorithmic_bytes(self):
        """Compulsory traffic of one y = A*x (SURVEY.md section 8d / BASELINE.md section 2)."""
        vb = self.val.element_size()
        return self.nnz * (vb + 4) + (self.rows + 1) * 4 + self.rows * vb + self.cols * vb


def uniform(rows, cols, nnz_per_row, dtype=torch.float64, values="ones", seed=0x5EED0001, device="cpu"):
    ro = uniform_row_offsets(rows, nnz_per_row)
    col, val = fill_nonzeros(ro, cols, 0, rows * nnz_per_row, kind="stratified", dtype=dtype,
                             values=values, seed=seed, device=device)
    return Csr(rows, cols, ro.to(device), col, val, f"uniform_{rows}x{cols}_{nnz_per_row}")


def powerlaw(rows, cols, max_row, target_nnz, dtype=torch.float32, values="ones", seed=0x5EED0003,
             device="cpu"):
    lengths, alpha = powerlaw_row_lengths(rows, min(max_row, cols), target_nnz, seed)
    ro = _offsets_from_lengths(lengths)
    col, val = fill_nonzeros(ro, cols, 0, int(ro[-1]), kind="stratified", dtype=dtype, values=values,
                             seed=seed, device=device)
    m = Csr(rows, cols, ro.to(device), col, val, f"powerlaw_{rows}_max{max_row}")
    m.alpha = alpha
    return m


def banded(rows, half_bandwidth=3, dtype=torch.float64, values="ones", seed=0x5EED0004, device="cpu"):
    ro = banded_row_offsets(rows, half_bandwidth)
    col, val = fill_nonzeros(ro, rows, 0, int(ro[-1]), kind="banded", dtype=dtype, values=values,
                             seed=seed, half_bandwidth=half_bandwidth, device=device)
    return Csr(rows, rows, ro.to(device), col, val, f"banded_{rows}_bw{2 * half_bandwidth + 1}")


def from_coo(rows, cols, r, c, v, dtype=torch.float64):
    """COO -> CSR the way CsrMatrix::Init does (sparse_matrix.h:666-728): stable sort by
    (row, col), duplicates kept, empty rows get repeated offsets."""
    r = np.asarray(r, dtype=np.int64)
    c = np.asarray(c, dtype=np.int64)
    v = np.asarray(v, dtype=np.float64)
    order = np.lexsort((c, r))  # stable: primary row, secondary col
    r, c, v = r[order], c[order], v[order]
    ro = np.zeros(rows + 1, dtype=np.int64)
    np.add.at(ro, r + 1, 1)
    ro = np.cumsum(ro)
    return Csr(rows, cols, torch.from_numpy(ro.astype(np.int32)), torch.from_numpy(c.astype(np.int32)),
               torch.from_numpy(v).to(dtype))


# ---------------------------------------------------------------------------------------------
# the reference's generators (values = 1.0): same edge order, then its stable sort
# ---------------------------------------------------------------------------------------------
def grid2d(width, dtype=torch.float64):
    """CooMatrix::InitGrid2d(width, self_loop=false) -- sparse_matrix.h:461-526 (W,E,N,S)."""
    n = width * width
    j, k = np.meshgrid(np.arange(width), np.arange(width), indexing="ij")
    me = (j * width + k).ravel()
    j, k = j.ravel(), k.ravel()
    r, c = [], []
    for mask, nb in ((k - 1 >= 0, me - 1), (k + 1 < width, me + 1), (j - 1 >= 0, me - width),
                     (j + 1 < width, me + width)):
        r.append(me[mask])
        c.append(nb[mask])
    r, c = np.concatenate(r), np.concatenate(c)
    m = from_coo(n, n, r, c, np.ones(r.size), dtype)
    m.name = f"grid2d_{width}"
    return m


def grid3d(width, dtype=torch.float64):
    """CooMatrix::InitGrid3d(width, self_loop=false) -- sparse_matrix.h:533-617."""
    n = width ** 3
    i, j, k = np.meshgrid(np.arange(width), np.arange(width), np.arange(width), indexing="ij")
    me = (i * width * width + j * width + k).ravel()
    i, j, k = i.ravel(), j.ravel(), k.ravel()
    r, c = [], []
    w2 = width * width
    for mask, nb in ((k - 1 >= 0, me - 1), (k + 1 < width, me + 1), (j - 1 >= 0, me - width),
                     (j + 1 < width, me + width), (i - 1 >= 0, me - w2), (i + 1 < width, me + w2)):
        r.append(me[mask])
        c.append(nb[mask])
    r, c = np.concatenate(r), np.concatenate(c)
    m = from_coo(n, n, r, c, np.ones(r.size), dtype)
    m.name = f"grid3d_{width}"
    return m


def wheel(spokes, dtype=torch.float64):
    """CooMatrix::InitWheel -- sparse_matrix.h:419-452: hub row 0 -> 1..spokes, rim i+1 -> ((i+1) % spokes)+1."""
    i = np.arange(spokes)
    r = np.concatenate([np.zeros(spokes, np.int64), i + 1])
    c = np.concatenate([i + 1, (i + 1) % spokes + 1])
    m = from_coo(spokes + 1, spokes + 1, r, c, np.ones(r.size), dtype)
    m.name = f"wheel_{spokes}"
    return m


def dense(rows, cols, dtype=torch.float64):
    """CooMatrix::InitDense -- sparse_matrix.h:386-413."""
    ro = (torch.arange(rows + 1, dtype=torch.int64) * cols).to(torch.int32)
    col = torch.arange(cols, dtype=torch.int32).repeat(rows)
    return Csr(rows, cols, ro, col, torch.ones(rows * cols, dtype=dtype), f"dense_{rows}x{cols}")


# ---------------------------------------------------------------------------------------------
# BASELINE.json configs by name
# ---------------------------------------------------------------------------------------------
CONFIGS = {
    # name: (kind, dtype, params)
    "cpu_uniform_16k": ("uniform", torch.float64, dict(rows=16384, cols=16384, nnz_per_row=32)),
    "uniform_1m_64": ("uniform", torch.float64, dict(rows=1 << 20, cols=1 << 20, nnz_per_row=64)),
    # same shape, columns drawn inside a 4096-wide window around the diagonal: shows the kernel
    # without the random-gather ceiling (SURVEY.md section 8d asks for a local-column variant)
    "uniform_1m_64_local": ("uniform_local", torch.float64, dict(rows=1 << 20, cols=1 << 20, nnz_per_row=64)),
    "powerlaw_2m": ("powerlaw", torch.float32, dict(rows=2_000_000, cols=2_000_000, max_row=1_000_000,
                                                   target_nnz=200_000_000)),
    "banded_10m": ("banded", torch.float64, dict(rows=10_000_000, half_bandwidth=3)),
    "powerlaw_20m": ("powerlaw", torch.float32, dict(rows=20_000_000, cols=20_000_000,
                                                    max_row=1_000_000, target_nnz=1_000_000_000)),
}


def make_config(name, values="ones", device="cpu", scale=1.0, dtype=None):
    """Build a BASELINE config (optionally scaled down by `scale` in rows/nnz for CPU tests)."""
    kind, dt, p = CONFIGS[name]
    dt = dtype or dt
    p = dict(p)
    if scale != 1.0:
        for key in ("rows", "cols", "target_nnz", "max_row"):
            if key in p:
                p[key] = max(8, int(p[key] * scale))
    if kind == "uniform":
        return uniform(dtype=dt, values=values, device=device, **p)
    if kind == "powerlaw":
        return powerlaw(dtype=dt, values=values, device=device, **p)
    return banded(dtype=dt, values=values, device=device, **p)
