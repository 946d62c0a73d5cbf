"""CPU tier (+ one GPU case): the multi-GPU row partitioning logic.  Concatenated per-rank row blocks must
reproduce the global CSR, and the partitioned aggregation (halo all-gather + local aggregation) must
equal the single-device result; exercised with 2 gloo ranks on CPU, with the oracle standing in for
the local aggregation (the CUDA path is covered by the GPU tier)."""
import os
import socket
import sys

import numpy as np
import pytest

from gnnagg import partition, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_split_rows_and_blocks_reassemble():
    ptr, idx = synth.small_random_csr(1000, 8.0, 1, hub=3000)
    val = np.random.default_rng(0).standard_normal(len(idx)).astype(np.float32)
    for parts in (1, 2, 3, 8):
        for balance in ("edges", "rows"):
            b = partition.split_rows(ptr, parts, balance)
            assert b[0] == 0 and b[-1] == 1000 and np.all(np.diff(b) >= 0) and len(b) == parts + 1
            blocks = [partition.local_block(ptr, idx, val, b, p) for p in range(parts)]
            assert np.array_equal(np.concatenate([x[1] for x in blocks]), idx)
            assert np.array_equal(np.concatenate([x[2] for x in blocks]), val)
            deg = np.concatenate([np.diff(x[0]) for x in blocks])
            assert np.array_equal(deg, np.diff(ptr))
            assert all(x[0][0] == 0 and x[0][-1] == len(x[1]) for x in blocks)
    b = partition.split_rows(ptr, 4, "edges")
    e = [int(ptr[b[p + 1]] - ptr[b[p]]) for p in range(4)]
    assert max(e) <= len(idx) / 4 + 3000 + 64  # balanced up to one (hub) row


def _worker(rank, world, port, balance, q):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    import oracle as orc

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ptr, idx = synth.small_random_csr(600, 6.0, 7, hub=1500)
        rng = np.random.default_rng(1)
        val = rng.standard_normal(len(idx)).astype(np.float32)
        X = rng.standard_normal((600, 32)).astype(np.float32)
        b = partition.split_rows(ptr, world, balance)
        lp, li, lv = partition.local_block(ptr, idx, val, b, rank)
        shard = torch.from_numpy(X[int(b[rank]):int(b[rank + 1])].copy())
        full = partition.all_gather_rows(shard, b)            # the halo exchange
        assert np.array_equal(full.numpy(), X)
        y_local, _ = orc.spmm_f64(lp, li, lv, full.numpy())   # local aggregation on global source ids
        y_ref, _ = orc.spmm_f64(ptr, idx, val, X)
        ok = np.array_equal(y_local, y_ref[int(b[rank]):int(b[rank + 1])])
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("balance", ["rows", "edges"])
def test_partitioned_equals_single_gloo_world2(balance):
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, balance, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=5) for _ in range(2))
    assert got == [(0, True), (1, True)]


@pytest.mark.gpu
def test_partitioned_blocks_on_gpu_equal_single(gn, orc, cuda):
    """every rank's block, run through the CUDA path against a replicated X, reproduces the 1-GPU rows"""
    import torch

    ptr, idx = synth.small_random_csr(3000, 10.0, 3, hub=20000)
    rng = np.random.default_rng(2)
    val = rng.standard_normal(len(idx)).astype(np.float32)
    X = rng.standard_normal((3000, 64)).astype(np.float32)
    dX = torch.from_numpy(X).to(cuda)
    single = gn.Aggregator(*(torch.from_numpy(a).to(cuda) for a in (ptr, idx, val)))
    Y1 = single.gcn_run(dX, torch.empty((3000, 64), device=cuda))
    for parts in (2, 4, 8):
        b = partition.split_rows(ptr, parts, "edges")
        for p in range(parts):
            lp, li, lv = partition.local_block(ptr, idx, val, b, p)
            rows = len(lp) - 1
            agg = gn.Aggregator(*(torch.from_numpy(a).to(cuda) for a in (lp, li, lv)))
            Yp = agg.gcn_run(dX, torch.empty((rows, 64), device=cuda))
            ref = Y1[int(b[p]):int(b[p + 1])]
            y64, scale = orc.spmm_f64(lp, li, lv, X)
            err = np.abs(Yp.cpu().numpy().astype(np.float64) - y64)
            assert np.all(err <= 1e-5 * scale + 1e-30)
            # same edges, possibly different item boundaries: equal within the same bound
            assert np.all(np.abs((Yp - ref).cpu().numpy()) <= 2e-5 * scale + 1e-30)
