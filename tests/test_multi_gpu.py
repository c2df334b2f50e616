"""N>1 on real GPUs: one process per GPU, both carry exchanges (NCCL all_gather + fold, NVLink peer-memory
kernel), eagerly and from CUDA graphs (skipped on boxes with a single GPU; the host logic of the same path is
covered on CPU by tests/test_sharded_gloo.py, the exchange kernel by tests/test_pipe_emu.py; last run on 2 and 8
GPUs: profiles/mg_sweep_r02_n*.txt)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, dtype_name, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import oracle
    import merge_spmv_b200 as ms
    from merge_spmv_b200 import generators as gen
    from merge_spmv_b200 import sharded

    dt = getattr(torch, dtype_name)
    m = gen.make_config("powerlaw_2m", scale=1 / 32, dtype=dt, values="random")
    ro, col, val = m.numpy()
    x = gen.vector(m.cols, dt, "random", device=dev)
    md = m.to(dev)
    shard = sharded.make_shard(ro, m.cols, rank, world, lambda k0, k1: (md.col[k0:k1], md.val[k0:k1]), dev)
    op = sharded.ShardedSpmv(shard)
    for _ in range(3):
        y_own = op(x)
    torch.cuda.synchronize()
    full = ms.csrmv(md.row_offsets, md.col, md.val, x)
    ok_local = torch.allclose(y_own, full[shard.x0:shard.x1], rtol=1e-5 if dt == torch.float32 else 1e-12, atol=0)
    want = oracle.Oracle().merge_csrmv(ro, col, val, x.cpu().numpy(), world)[shard.x0:shard.x1]
    got = y_own.cpu().numpy()
    lens = np.diff(ro)[shard.x0:shard.x1].astype(np.float64)
    tol = 1e-10 if dt == torch.float64 else np.maximum(1e-6, 4 * np.sqrt(lens) * 2.0 ** -24)
    ok_oracle = bool(np.all(np.abs(got - want) <= tol * np.abs(want)))
    # the NVLink peer-memory carry exchange (bench.py's default) instead of NCCL: same carries, same fold
    # order, so the same bits -- eagerly and replayed from a CUDA graph
    op2 = sharded.ShardedSpmv(shard, exchange="p2p")
    for _ in range(4):  # several epochs: both parities of the double-buffered slots
        y2 = op2(x)
    torch.cuda.synchronize()
    ok_p2p = torch.equal(y2, y_own)
    rep2 = op2.capture(x)
    for _ in range(3):
        y3 = rep2()
    torch.cuda.synchronize()
    ok_p2p = ok_p2p and torch.equal(y3, y_own)
    del rep2
    # the y exchange (every rank gets the whole vector), eagerly and replayed from a CUDA graph
    rtol = 1e-5 if dt == torch.float32 else 1e-12
    ok_full = torch.allclose(op.matvec_full(x), full, rtol=rtol, atol=0)
    replay = op.capture(x, gather_y=True)  # the y exchange replayed from a CUDA graph
    ok_full = ok_full and torch.allclose(replay(), full, rtol=rtol, atol=0)
    torch.cuda.synchronize()
    flag = torch.tensor([int(ok_local and ok_oracle and ok_full and ok_p2p)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        np.save(os.path.join(out_dir, f"ok_{dtype_name}.npy"), np.array([int(flag.item())]))
    del replay  # a captured graph holds NCCL work: release it before the communicator
    torch.cuda.synchronize()
    dist.destroy_process_group()


@pytest.mark.parametrize("dtype_name", ["float64", "float32"])
def test_sharded_spmv_nccl(tmp_path, dtype_name):
    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(world, 8)
    mp.spawn(_worker, args=(world, _free_port(), dtype_name, str(tmp_path)), nprocs=world, join=True)
    assert np.load(tmp_path / f"ok_{dtype_name}.npy")[0] == 1
