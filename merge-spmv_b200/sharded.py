"""One-process-per-GPU sharding of a single CsrMV by the merge decomposition.

New surface (the reference is single-GPU; README.md:5 and the paper's section III.A only state
that the decomposition partitions hierarchically).  The cut is the one ``OmpMergeCsrmv`` makes
between CPU threads (cpu_spmv.cpp:311-321) with p = world size:

* rank g owns diagonals ``[g*ceil((rows+nnz)/p), ...)`` of the merge path, i.e. nonzeros
  ``[y_g, y_{g+1})`` and the rows ``[x_g, x_{g+1})`` that END inside its span;
* its shard is an ordinary local CSR with ``x_{g+1} - x_g + 1`` rows -- the last local row is the
  (possibly empty) leading part of global row ``x_{g+1}`` -- so the same ``mspmv_csrmv_*`` kernel
  runs unchanged and ``y_local[-1]`` is the carry-out (cpu_spmv.cpp:336-344);
* carry *rows* are known to every rank from the partition, so only the p carry *values* are
  exchanged: ONE ``all_gather`` (NCCL over NVLink) straight out of ``y_local[-1:]``;
* ``mspmv_apply_carries_*`` then folds carries 0..p-2 into the owning rank's slice in shard order
  with the ``row < num_rows`` guard -- the serial fix-up of cpu_spmv.cpp:348-352.

``x`` is replicated, ``y`` stays sharded (rank g holds global rows ``[x_g, x_{g+1})``).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .csrmv import _ptr, _stream, csrmv, temp_storage


def partition(row_offsets_host: np.ndarray, world: int) -> np.ndarray:
    """(world+1, 2) int32 cut coordinates (row, nonzero); host-side MergePathSearch."""
    ro = np.ascontiguousarray(row_offsets_host, dtype=np.int32)
    rows, nnz = ro.size - 1, int(ro[-1])
    coords = np.zeros((world + 1, 2), dtype=np.int32)
    _lib.lib().mspmv_shard_partition(ro.ctypes.data_as(C.c_void_p), rows, nnz, world,
                                     coords.ctypes.data_as(C.c_void_p))
    return coords


def local_row_offsets(row_offsets_host: np.ndarray, coords: np.ndarray, rank: int) -> np.ndarray:
    ro = np.ascontiguousarray(row_offsets_host, dtype=np.int32)
    (x0, y0), (x1, y1) = coords[rank], coords[rank + 1]
    out = np.zeros(int(x1 - x0) + 2, dtype=np.int32)
    _lib.lib().mspmv_shard_row_offsets(ro.ctypes.data_as(C.c_void_p), int(x0), int(y0), int(x1), int(y1),
                                       out.ctypes.data_as(C.c_void_p))
    return out


@dataclass
class Shard:
    rank: int
    world: int
    coords: np.ndarray          # (world+1, 2)
    rows_global: int
    cols: int
    row_offsets: torch.Tensor   # int32 [local_rows + 1], rebased to this shard's first nonzero
    col: torch.Tensor           # int32 [y1 - y0]
    val: torch.Tensor           # [y1 - y0]
    carry_rows: torch.Tensor    # int32 [world]: global row each rank's carry belongs to

    @property
    def x0(self):
        return int(self.coords[self.rank, 0])

    @property
    def x1(self):
        return int(self.coords[self.rank + 1, 0])

    @property
    def y0(self):
        return int(self.coords[self.rank, 1])

    @property
    def y1(self):
        return int(self.coords[self.rank + 1, 1])

    @property
    def owned_rows(self):
        return self.x1 - self.x0

    @property
    def local_rows(self):
        return self.owned_rows + 1

    @property
    def nnz(self):
        return self.y1 - self.y0


def make_shard(row_offsets_host, cols, rank, world, fill, device) -> Shard:
    """Build rank's shard.  ``fill(k0, k1)`` returns (col int32, val) tensors on ``device`` for the
    global nonzero range [k0, k1) -- a slice of resident arrays or a generator."""
    ro = np.ascontiguousarray(row_offsets_host, dtype=np.int32)
    coords = partition(ro, world)
    lro = local_row_offsets(ro, coords, rank)
    y0, y1 = int(coords[rank, 1]), int(coords[rank + 1, 1])
    col, val = fill(y0, y1)
    # slices of resident arrays start at arbitrary elements; give the kernels 16-byte-aligned bases
    # (they accept misaligned ones, but then stage ragged edges with scalar copies)
    if col.data_ptr() % 16:
        col = col.clone()
    if val.data_ptr() % 16:
        val = val.clone()
    return Shard(rank=rank, world=world, coords=coords, rows_global=ro.size - 1, cols=int(cols),
                 row_offsets=torch.from_numpy(lro).to(device), col=col.contiguous(), val=val.contiguous(),
                 carry_rows=torch.from_numpy(np.ascontiguousarray(coords[1:, 0])).to(device))


def apply_carries(y_local, shard: Shard, carry_vals, stream=None):
    """Device fold of the gathered carries into this rank's owned rows (cpu_spmv.cpp:348-352)."""
    sfx = "f64" if y_local.dtype == torch.float64 else "f32"
    fn = getattr(_lib.lib(), f"mspmv_apply_carries_{sfx}")
    with torch.cuda.device(y_local.device):
        _lib.check(fn(_ptr(y_local), shard.x0, shard.owned_rows, shard.rows_global,
                      _ptr(shard.carry_rows), _ptr(carry_vals), shard.world, _stream(stream)),
                   "apply_carries")


class ShardedSpmv:
    """y_shard = (A x)[x_g : x_{g+1}) on this rank; call collectively on all ranks."""

    def __init__(self, shard: Shard, group=None, local_spmv=None, fold=None, exchange="nccl"):
        """``exchange``: how the p carry values travel.  "nccl": one ``all_gather`` + the fold kernel
        (the portable form, and what the CPU / gloo tests exercise).  "p2p": ONE kernel per rank that
        stores its carry straight into every peer's symmetric-memory buffer over NVLink, waits for the
        peers' flags and folds (``mspmv_exchange_carries_*``, csrc/carry_exchange.cuh) -- no NCCL on the
        data path; measured 0.330 vs 0.409 ms per step at 8 GPUs and a tie at 2
        (profiles/mg_sweep_r02_n*.txt), so bench.py and anything latency-bound should pass "p2p"."""
        if exchange not in ("nccl", "p2p"):
            raise ValueError("exchange must be 'nccl' or 'p2p'")
        self.shard = shard
        self.group = group
        self._exchange = exchange
        dev = shard.val.device
        # y slices have different lengths (equal WORK per rank, not equal rows): the exchange of y
        # (gather_y) sends fixed-size records of `pad` rows, so the local result lives at the head
        # of a buffer of at least that size and is sent in place.
        self._sizes = [int(shard.coords[g + 1, 0] - shard.coords[g, 0]) for g in range(shard.world)]
        self._pad = max(max(self._sizes), 1)
        self._ybuf = torch.zeros(max(shard.local_rows, self._pad), dtype=shard.val.dtype, device=dev)
        self.y_local = self._ybuf[:shard.local_rows]
        self.carries = torch.zeros(shard.world, dtype=shard.val.dtype, device=dev)
        self._gathered = None  # (world, pad) receive buffer of gather_y, allocated on first use
        # injectable for the CPU (gloo) tests of the exchange logic; the defaults are the CUDA path
        # own temp blob: a captured graph bakes its address in, so it must not be a shared, replaceable one
        self._temp = None
        if local_spmv is None:
            self._temp = temp_storage(shard.val.dtype, shard.local_rows, shard.cols, shard.nnz, dev)
        self._local_spmv = local_spmv or (lambda s, x, y: csrmv(s.row_offsets, s.col, s.val, x, y,
                                                               num_cols=s.cols, temp=self._temp))
        self._fold = fold or apply_carries
        if exchange == "p2p" and shard.world > 1:
            self._setup_p2p()

    def _setup_p2p(self):
        """Symmetric exchange buffers: 4*p 64-bit words per rank (values[2][p], flags[2][p]), mapped
        into every process of the group (torch.distributed._symmetric_memory: cudaIpc / VMM)."""
        import torch.distributed._symmetric_memory as symm_mem

        s = self.shard
        dev = self.y_local.device
        if dev.type != "cuda":
            raise _lib.MergeSpmvError("the p2p carry exchange needs CUDA devices with peer access (NVLink)")
        group = self.group if self.group is not None else dist.group.WORLD
        nbytes = _lib.lib().mspmv_exchange_buffer_bytes(s.world)
        try:
            symm_mem.enable_symm_mem_for_group(group.group_name)
        except Exception:
            pass  # newer torch enables it on rendezvous
        self._xbuf = symm_mem.empty(nbytes // 8, dtype=torch.int64, device=dev)
        self._xbuf.zero_()
        self._xhdl = symm_mem.rendezvous(self._xbuf, group.group_name)
        if self._xhdl.world_size != s.world or self._xhdl.rank != s.rank:
            raise _lib.MergeSpmvError("symmetric-memory group does not match the shard layout")
        # Address of the buffer in every rank's mapping.  The handle reports the base of each rank's
        # symmetric allocation; the tensor may sit at an offset inside it (pool allocators), which is
        # the same on every rank.  Check the arithmetic against the one address we know: our own.
        bases = [int(p) for p in self._xhdl.buffer_ptrs]
        off = int(getattr(self._xhdl, "offset", 0) or 0)
        if bases[s.rank] + off != self._xbuf.data_ptr():
            off = self._xbuf.data_ptr() - bases[s.rank]
            if off < 0 or off + nbytes > int(self._xhdl.buffer_size):
                raise _lib.MergeSpmvError("cannot locate the exchange buffer inside the symmetric allocation")
        self._peer_table = torch.tensor([b + off for b in bases], dtype=torch.int64, device=dev)
        torch.cuda.synchronize(dev)
        self._xhdl.barrier()  # every buffer is zero before anybody pushes
        self._epoch = torch.zeros(1, dtype=torch.int64, device=dev)

    def _exchange_p2p(self):
        s = self.shard
        sfx = "f64" if self.y_local.dtype == torch.float64 else "f32"
        fn = getattr(_lib.lib(), f"mspmv_exchange_carries_{sfx}")
        with torch.cuda.device(self.y_local.device):
            _lib.check(fn(_ptr(self.y_local), s.local_rows, s.x0, s.owned_rows, s.rows_global,
                          _ptr(s.carry_rows), _ptr(self._peer_table), s.rank, s.world,
                          _ptr(self._epoch), _stream(None)), "exchange_carries")

    def capture(self, x, gather_y=False):
        """Record one whole step (the CsrMV kernel, the all_gather, the carry fold --
        and, with ``gather_y``, the exchange of the y slices) into a CUDA graph over the fixed
        input buffer ``x``; returns a callable that replays it with a single launch.  Removes
        per-kernel launch gaps and the host cost of the collectives."""
        step = self.matvec_full if gather_y else self
        side = torch.cuda.Stream(device=self.y_local.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):  # warm-up outside capture: one-time attribute setup, NCCL channels
                step(x)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = step(x)

        def replay():
            graph.replay()
            return out

        replay.graph = graph
        return replay

    def gather_y(self, out=None):
        """The step on the far side of the path for iterative solvers: after a
        product every rank holds only its rows of y, but the next product needs the whole vector
        as x.  ONE ``all_gather`` of fixed-size records (``pad`` = longest slice) taken in place
        from the local result, then ``world`` contiguous copies drop the padding.  Call
        collectively, after ``self(x)``.  Returns the full y (length ``rows_global``)."""
        s = self.shard
        dt, dev = self.y_local.dtype, self.y_local.device
        if out is None:
            out = torch.empty(s.rows_global, dtype=dt, device=dev)
        if s.world == 1:
            out.copy_(self.y_local[:s.owned_rows])
            return out
        if self._gathered is None:
            self._gathered = torch.empty(s.world * self._pad, dtype=dt, device=dev)
        dist.all_gather_into_tensor(self._gathered, self._ybuf[:self._pad], group=self.group)
        for g, n in enumerate(self._sizes):
            if n:
                x_g = int(s.coords[g, 0])
                out[x_g:x_g + n].copy_(self._gathered[g * self._pad:g * self._pad + n])
        return out

    def matvec_full(self, x, out=None):
        """``y = A x`` replicated on every rank: the sharded product followed by ``gather_y``."""
        self(x)
        return self.gather_y(out)

    def __call__(self, x):
        s = self.shard
        self._local_spmv(s, x, self.y_local)
        if s.world > 1 and self._exchange == "p2p":
            self._exchange_p2p()  # push over NVLink + wait + fold, one launch
        elif s.world > 1:
            # the single exchange step: p carry values, taken in place from the last local row
            dist.all_gather_into_tensor(self.carries, self.y_local[s.local_rows - 1:], group=self.group)
            self._fold(self.y_local, s, self.carries)
        return self.y_local[:s.owned_rows]
