"""The N>1 path on CPU: world_size-2 (and 3) gloo process groups exercise the host-side sharding
logic -- partition, local CSR construction, the single all_gather of carry values, ownership of
rows -- with the local SpMV and the fold injected from the oracle (tests may use the oracle as the
checker; the product's CUDA path is covered by tests/test_gpu_parity.py and bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, dtype_name, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from merge_spmv_b200 import generators as gen
    from merge_spmv_b200 import sharded

    orc = oracle.Oracle()
    dt = getattr(torch, dtype_name)
    m = gen.make_config("powerlaw_2m", scale=1 / 200, dtype=dt, values="random")
    ro, col, val = m.numpy()
    x = gen.vector(m.cols, dt, "random")

    def local_spmv(s, xv, y_out):
        y = orc.merge_csrmv(s.row_offsets.numpy(), s.col.numpy(), s.val.numpy(), xv.numpy(), 1)
        y_out.copy_(torch.from_numpy(y))

    def fold(y_local, s, carries):  # cpu_spmv.cpp:348-352
        rows = s.carry_rows.numpy()
        for g in range(s.world - 1):
            r = int(rows[g])
            if r < s.rows_global and s.x0 <= r < s.x1:
                y_local[r - s.x0] += carries[g]

    shard = sharded.make_shard(ro, m.cols, rank, world, lambda k0, k1: (m.col[k0:k1], m.val[k0:k1]), "cpu")
    op = sharded.ShardedSpmv(shard, local_spmv=local_spmv, fold=fold)
    y_own = op(x).clone()

    # gather the y slices (test-only) and compare with the oracle run with p = world threads
    sizes = [int(shard.coords[g + 1, 0] - shard.coords[g, 0]) for g in range(world)]
    assert sum(sizes) == m.rows
    pad = max(sizes)  # gloo all_gather needs equal sizes
    mine = torch.zeros(pad, dtype=dt)
    mine[:y_own.numel()] = y_own
    pieces = [torch.empty(pad, dtype=dt) for _ in sizes]
    dist.all_gather(pieces, mine)
    want = orc.merge_csrmv(ro, col, val, x.numpy(), world)
    got = torch.cat([p[:n] for p, n in zip(pieces, sizes)]).numpy()

    # the exchange on the far side of the path (solver-style ping-pong): every rank gets the whole y
    # from ONE all_gather of padded slices, feeds it back as the next x
    y_full = op.matvec_full(x).clone()
    ok_full = np.array_equal(y_full.numpy(), want)
    x2 = y_full / y_full.abs().max()
    y2 = op.matvec_full(x2)
    want2 = orc.merge_csrmv(ro, col, val, x2.numpy(), world)
    ok_chain = np.array_equal(y2.numpy(), want2)
    flag = torch.tensor([int(ok_full and ok_chain)])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        np.save(os.path.join(out_dir, f"ok_{dtype_name}_{world}.npy"),
                np.array([np.array_equal(got, want), got.size, int(flag.item())]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("dtype_name", ["float64", "float32"])
def test_sharded_spmv_gloo(tmp_path, world, dtype_name):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, dtype_name, str(tmp_path)), nprocs=world, join=True)
    res = np.load(tmp_path / f"ok_{dtype_name}_{world}.npy")
    assert res[0] == 1, "sharded result differs from the oracle with p = world threads"
    assert res[1] == 10000
    assert res[2] == 1, "gather_y / matvec_full (replicated y, fed back as x) differs from the oracle on some rank"
