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


def _worker_pruned(rank, world, port, q):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    import oracle as orc

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_per, F = 200, 16
        rng = np.random.default_rng(3)                       # same stream on every rank: the global X
        Xfull = rng.standard_normal((world * n_per, F)).astype(np.float32)
        ptr, idx = synth.small_random_csr(n_per, 5.0, 40 + rank, num_src=world * n_per, hub=400)
        # make the referenced set sparse so pruning matters: keep only sources = 0 mod 3
        idx = (idx // 3 * 3).astype(np.int32)
        val = np.random.default_rng(rank).standard_normal(len(idx)).astype(np.float32)
        plan = partition.pruned_plan(torch.from_numpy(idx), n_per, world, rank)
        shard = torch.from_numpy(Xfull[rank * n_per:(rank + 1) * n_per].copy())
        send = shard[plan["send_rows"]]                       # the packing step (gnnagg_gather_rows on the GPU)
        recv = torch.empty((plan["num_recv"], F))
        dist.all_to_all_single(recv, send, output_split_sizes=plan["recv_counts"], input_split_sizes=plan["send_counts"])
        y, _ = orc.spmm_f64(ptr, plan["idx_compact"].numpy(), val, recv.numpy())
        want, _ = orc.spmm_f64(ptr, idx, val, Xfull)
        ok = np.array_equal(y, want) and plan["num_recv"] <= world * n_per // 3 + 1
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def _worker_incremental(rank, world, port, q):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    import oracle as orc

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_per, F, chunks = 300, 8, 3
        Xfull = np.random.default_rng(3).standard_normal((world * n_per, F)).astype(np.float32)
        ptr, idx = synth.small_random_csr(n_per, 6.0, 60 + rank, num_src=world * n_per, hub=700)
        idx = (idx // 2 * 2).astype(np.int32)
        val = np.random.default_rng(rank).standard_normal(len(idx)).astype(np.float32)
        plan = partition.incremental_plan(torch.from_numpy(ptr), torch.from_numpy(idx), n_per, world, rank, chunks)
        shard = torch.from_numpy(Xfull[rank * n_per:(rank + 1) * n_per].copy())
        recv = torch.full((plan["num_recv"], F), float("nan"))
        Y = np.zeros((n_per, F), np.float32)
        rb = plan["row_bounds"]
        ok = True
        for c in range(chunks):
            o, cnt = plan["recv_offset"][c], sum(plan["recv_counts"][c])
            send = shard[plan["send_rows"][c]]
            out = torch.empty((cnt, F))
            dist.all_to_all_single(out, send, output_split_sizes=plan["recv_counts"][c], input_split_sizes=plan["send_counts"][c])
            recv[o:o + cnt] = out
            r0, r1 = int(rb[c]), int(rb[c + 1])
            e0, e1 = int(ptr[r0]), int(ptr[r1])
            sub_idx = plan["idx_compact"][e0:e1].numpy()
            # chunk c may only touch rows that have arrived by stage c
            ok = ok and (len(sub_idx) == 0 or sub_idx.max() < o + cnt)
            y, _ = orc.spmm_f64((ptr[r0:r1 + 1] - e0).astype(np.int32), sub_idx, val[e0:e1], recv.numpy())
            Y[r0:r1] = y
        want, _ = orc.spmm_f64(ptr, idx, val, Xfull)
        ok = ok and np.array_equal(Y, want) and plan["num_recv"] == len(np.unique(idx))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_incremental_halo_plan_gloo_world2():
    """pipelined pruned exchange: each row chunk fetches only sources no earlier chunk fetched; results equal"""
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_incremental, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert sorted(q.get(timeout=5) for _ in range(2)) == [(0, True), (1, True)]


def test_pruned_halo_plan_gloo_world2():
    """only the referenced source rows travel; the re-indexed CSR on the compact buffer gives identical sums"""
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_pruned, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert sorted(q.get(timeout=5) for _ in range(2)) == [(0, True), (1, True)]


@pytest.mark.gpu
def test_gather_rows_kernel(gn, cuda):
    import torch

    X = torch.randn((1000, 64), device=cuda)
    rows = torch.randint(0, 1000, (5000,), device=cuda, dtype=torch.int64)
    out = gn.gather_rows(X, rows, torch.empty((5000, 64), device=cuda))
    assert torch.equal(out, X[rows])
    assert gn.gather_rows(X, rows[:0], torch.empty((0, 64), device=cuda)).numel() == 0


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


def test_split_by_source_chunk_reassembles():
    """sub-CSRs of the pipelined halo exchange: every edge lands in exactly one sub-CSR and the remapped
    indices address the chunk buffers correctly (oracle as the aggregation)"""
    import torch

    import oracle as orc

    world, n_per, F, chunks = 3, 50, 8, 4
    rng = np.random.default_rng(5)
    Xfull = rng.standard_normal((world * n_per, F)).astype(np.float32)
    for rank in range(world):
        ptr, idx = synth.small_random_csr(n_per, 9.0, 20 + rank, num_src=world * n_per, hub=300)
        val = rng.standard_normal(len(idx)).astype(np.float32)
        parts = partition.split_by_source_chunk(torch.from_numpy(ptr), torch.from_numpy(idx), torch.from_numpy(val),
                                                n_per, world, rank, chunks)
        assert sum(p[1].numel() for p in parts) == len(idx)
        cb = partition.chunk_rows(n_per, chunks)
        acc = np.zeros((n_per, F), np.float64)
        Xs = Xfull[rank * n_per:(rank + 1) * n_per]
        y, _ = orc.spmm_f64(parts[0][0].numpy(), parts[0][1].numpy(), parts[0][2].numpy(), np.ascontiguousarray(Xs))
        acc += y
        for c in range(chunks):
            Xc = np.concatenate([Xfull[r * n_per + cb[c]: r * n_per + cb[c + 1]] for r in range(world)])  # all-gather c
            p, i, v = (t.numpy() for t in parts[1 + c])
            if len(i):
                assert i.max() < len(Xc)
            y, _ = orc.spmm_f64(p, i, v, np.ascontiguousarray(Xc))
            acc += y
        want, scale = orc.spmm_f64(ptr, idx, val, Xfull)
        assert np.all(np.abs(acc - want) <= 1e-5 * scale + 1e-30)


@pytest.mark.gpu
def test_accumulate_mode_sums_sub_csrs(gn, orc, cuda):
    """gnnagg_gcn_run_acc: Y = A0*X then Y += A1*X ... equals the un-split aggregation"""
    import torch

    world, n_per, F = 4, 700, 64
    rng = np.random.default_rng(9)
    ptr, idx = synth.small_random_csr(n_per, 30.0, 4, num_src=world * n_per, hub=5000)
    val = rng.standard_normal(len(idx)).astype(np.float32)
    Xfull = rng.standard_normal((world * n_per, F)).astype(np.float32)
    rank, chunks = 1, 3
    tp, ti, tv = (torch.from_numpy(a).to(cuda) for a in (ptr, idx, val))
    parts = partition.split_by_source_chunk(tp, ti, tv, n_per, world, rank, chunks)
    cb = partition.chunk_rows(n_per, chunks)
    Y = torch.full((n_per, F), float("nan"), device=cuda)
    dX = torch.from_numpy(Xfull).to(cuda)
    aggs = [gn.Aggregator(*p) for p in parts]
    aggs[0].gcn_run_acc(dX[rank * n_per:(rank + 1) * n_per].contiguous(), Y, accumulate=False)
    for c in range(chunks):
        Xc = torch.cat([dX[r * n_per + cb[c]: r * n_per + cb[c + 1]] for r in range(world)]).contiguous()
        aggs[1 + c].gcn_run_acc(Xc, Y, accumulate=True)
    want, scale = orc.spmm_f64(ptr, idx, val, Xfull)
    err = np.abs(Y.cpu().numpy().astype(np.float64) - want)
    assert np.all(err <= 1e-5 * scale + 1e-30)
