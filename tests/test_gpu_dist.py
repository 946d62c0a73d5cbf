"""GPU tier: the peer-memory multi-GPU path (gnnagg_dist_*, csrc/dist.cu) against the oracle.

On a one-GPU box all ranks live on device 0 (the peer pointers are then ordinary device pointers): the flag
protocol, the pull kernels, the device-built plan and the staged accumulation are exercised exactly as across
GPUs.  With >= 2 GPUs the torchrun test below runs one process per GPU over cudaIpc mappings / NVLink and
compares with the single-GPU result on a down-scaled R-MAT graph."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT
from gnnagg import partition, synth

pytestmark = pytest.mark.gpu


def _run_local(gn, cuda, ptr, idx, val, X, bounds, stages, layer_W=None, steps=1):
    world, F = len(bounds) - 1, X.shape[1]
    ld = partition.LocalDist(bounds, F, devices=[0] * world)
    streams = [torch.cuda.Stream(device=cuda) for _ in range(world)]
    dptr, didx, dval = (torch.from_numpy(a).to(cuda) for a in (ptr, idx, val))
    outs = []
    try:
        for r in range(world):
            lp, li, lv = partition.local_block(ptr, idx, val, bounds, r)
            ld.set_graph(r, torch.from_numpy(lp).to(cuda), torch.from_numpy(li).to(cuda), torch.from_numpy(lv).to(cuda), stages)
            plan = partition.peer_plan(li, bounds, r, stages, ptr=lp)
            ph = ld.ranks[r]
            assert ph.num_recv == len(plan["recv_rows"])
            assert ph.recv_counts == [int(c) for c in plan["recv_counts"]]
            assert ph.num_stages == plan["num_stages"]
            assert ph.stage_edges == [int((plan["stage"] == s).sum()) for s in range(plan["num_stages"])]
        ld.connect()
        for r in range(world):   # what every owner pushes is what the receivers asked for
            assert ld.ranks[r].send_counts == [ld.ranks[q].recv_counts[r] for q in range(world)]
        Wd = None if layer_W is None else torch.from_numpy(layer_W).to(cuda)
        for step in range(steps):
            buf = step & 1
            Xs = X if step == 0 else (X * (1.0 + step)).astype(np.float32)
            for r in range(world):
                ld.ranks[r].x(buf, F).copy_(torch.from_numpy(np.ascontiguousarray(Xs[bounds[r]:bounds[r + 1]])).to(cuda))
            torch.cuda.synchronize()
            Y = [torch.full((bounds[r + 1] - bounds[r], F if Wd is None else Wd.shape[1]), float("nan"), device=cuda) for r in range(world)]
            for r in range(world):
                with torch.cuda.stream(streams[r]):
                    if Wd is None:
                        ld.ranks[r].gcn_run(Y[r], buf, F)
                    else:
                        ld.ranks[r].gcn_layer(Wd, Y[r], buf)
            torch.cuda.synchronize()
            for r in range(world):
                ld.ranks[r].check()
            outs.append(torch.cat(Y).cpu().numpy())
    finally:
        ld.close()
    return outs


@pytest.mark.parametrize("world,stages", [(1, 1), (2, 0), (2, 1), (3, 1), (3, 2), (4, 0), (4, 3), (2, -4), (3, -16), (4, -3)])
@pytest.mark.parametrize("F", [32, 64, 128, 100])
def test_local_ranks_match_oracle(gn, orc, cuda, world, stages, F):
    rng = np.random.default_rng(world * 100 + stages * 10 + F)
    sizes = rng.integers(300, 900, world)
    bounds = [0] + [int(b) for b in np.cumsum(sizes)]
    n = bounds[-1]
    ptr, idx = synth.small_random_csr(n, 20.0, 5 + world, hub=9000)
    val = rng.standard_normal(len(idx)).astype(np.float32)
    X = rng.standard_normal((n, F)).astype(np.float32)
    want, scale = orc.spmm_f64(ptr, idx, val, X)
    outs = _run_local(gn, cuda, ptr, idx, val, X, bounds, stages, steps=2)
    for step, Y in enumerate(outs):
        err = np.abs(Y.astype(np.float64) - want * (1.0 + step))
        assert np.all(err <= 1e-5 * scale * (1.0 + step) + 1e-30), (step, float((err / (1e-5 * scale * (1 + step) + 1e-30)).max()))


def test_local_ranks_layer_and_determinism(gn, orc, cuda):
    rng = np.random.default_rng(3)
    bounds = [0, 700, 700, 1900, 2500]          # rank 1 owns nothing
    n, F = bounds[-1], 64
    ptr, idx = synth.small_random_csr(n, 30.0, 9, hub=20000)
    val = rng.standard_normal(len(idx)).astype(np.float32)
    X = rng.standard_normal((n, F)).astype(np.float32)
    W = (rng.standard_normal((F, 32)) / 8).astype(np.float32)
    _, h64, hs = orc.gcn_layer_f64(ptr, idx, val, X, W)
    a = _run_local(gn, cuda, ptr, idx, val, X, bounds, 2, layer_W=W)[0]
    b = _run_local(gn, cuda, ptr, idx, val, X, bounds, 2, layer_W=W)[0]
    assert np.all(np.abs(a.astype(np.float64) - h64) <= 1e-5 * hs + 1e-30)
    assert np.array_equal(a, b)                  # no atomics anywhere: bit-reproducible


@pytest.mark.parametrize("layer", [False, True])
def test_host_buffer_entry_point_single_rank(gn, orc, cuda, layer):
    """gnnagg_dist_gcn_layer_host on a one-rank group (it synchronises the stream, so several ranks need one process or
    host thread each: tests/multirank_worker.py covers that): row-chunked last stage + copy back vs the oracle"""
    rng = np.random.default_rng(11)
    n, F = 9000, 64
    ptr, idx = synth.small_random_csr(n, 25.0, 21, hub=30000)
    val = rng.standard_normal(len(idx)).astype(np.float32)
    X = rng.standard_normal((n, F)).astype(np.float32)
    W = (rng.standard_normal((F, 96)) / 8).astype(np.float32)
    ld = partition.LocalDist([0, n], F, devices=[0])
    try:
        ph = ld.set_graph(0, *(torch.from_numpy(a).to(cuda) for a in (ptr, idx, val)), 0)
        ld.connect()
        hX = torch.from_numpy(X).pin_memory()
        hW = torch.from_numpy(W).pin_memory() if layer else None
        hH = torch.empty((n, 96 if layer else F)).pin_memory()
        ph.gcn_layer_host(hX, hW, hH)
        ph.check()
    finally:
        ld.close()
    if layer:
        _, want, scale = orc.gcn_layer_f64(ptr, idx, val, X, W)
    else:
        want, scale = orc.spmm_f64(ptr, idx, val, X)
    assert np.all(np.abs(hH.numpy().astype(np.float64) - want) <= 1e-5 * scale + 1e-30)


def test_bad_source_id_is_rejected(gn, cuda):
    bounds = [0, 10, 20]
    ld = partition.LocalDist(bounds, 32, devices=[0, 0])
    try:
        ptr = torch.tensor([0] + [1] * 10, dtype=torch.int32, device=cuda)
        idx = torch.tensor([25], dtype=torch.int32, device=cuda)
        with pytest.raises(gn.GnnaggError):
            ld.set_graph(0, ptr, idx, torch.ones(1, device=cuda), 1)
        with pytest.raises(gn.GnnaggError):       # connecting ranks without a graph is a state error, not a crash
            ld.connect()
    finally:
        ld.close()


def test_cpp_caller_dist_check():
    """tests/cpp/dist_check.cu: a C++ caller of gnnagg_dist_create / _set_graph / _gcn_run / _gcn_layer"""
    exe = os.path.join(ROOT, "build", "compat", "dist_check.out")
    if not os.path.exists(exe):
        pytest.fail("build/compat/dist_check.out missing: run __graft_entry__.build()")
    for argv in ([], ["4", "2000", "16", "128", "3"]):
        if torch.cuda.device_count() > 1 and argv:
            argv = [str(min(8, torch.cuda.device_count()))] + argv[1:]
        r = subprocess.run([exe] + argv, capture_output=True, text=True, timeout=600)
        print(r.stdout[-2000:], r.stderr[-2000:])
        assert r.returncode == 0 and "DIST_CHECK ok" in r.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs (run under gpurun --gpus N)")
def test_multi_process_equals_single_gpu():
    """one process per GPU (torchrun, cudaIpc peer mappings, NVLink): the partitioned RMAT-22 aggregation equals the
    single-GPU result; log kept in gpurun_out/"""
    nproc = min(8, torch.cuda.device_count())
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "tests", "multirank_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    with open(os.path.join(out_dir, "multirank_n%d.log" % nproc), "w") as f:
        f.write(r.stdout + "\n--- stderr ---\n" + r.stderr[-4000:])
    print(r.stdout[-3000:], r.stderr[-2000:])
    assert r.returncode == 0
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    assert lines and all(l["ok"] for l in lines)


def test_narrower_feature_width_than_feat_cap_and_wide_rows(gn, orc, cuda):
    """feat < feat_cap re-uses the same peer-visible buffers with a tighter row stride (every rank is prepared for the new
    width before the first step: a process that drives several ranks must do that, gnnagg_dist_prepare); F = 256 takes
    the two-register-chunk kernel variant"""
    rng = np.random.default_rng(41)
    bounds = [0, 500, 1300, 2000]
    n, cap = bounds[-1], 256
    ptr, idx = synth.small_random_csr(n, 15.0, 17, hub=7000)
    val = rng.standard_normal(len(idx)).astype(np.float32)
    ld = partition.LocalDist(bounds, cap, devices=[0, 0, 0])
    streams = [torch.cuda.Stream(device=cuda) for _ in range(3)]
    try:
        for r in range(3):
            lp, li, lv = partition.local_block(ptr, idx, val, bounds, r)
            ld.set_graph(r, torch.from_numpy(lp).to(cuda), torch.from_numpy(li).to(cuda), torch.from_numpy(lv).to(cuda), 2)
        ld.connect()
        for F in (256, 64):
            X = rng.standard_normal((n, F)).astype(np.float32)
            want, scale = orc.spmm_f64(ptr, idx, val, X)
            for r in range(3):
                gn.check(gn.lib().gnnagg_dist_prepare(ld.ranks[r].h, F, None))
                ld.ranks[r].x(0, F).copy_(torch.from_numpy(np.ascontiguousarray(X[bounds[r]:bounds[r + 1]])).to(cuda))
            torch.cuda.synchronize()
            Y = [torch.empty((bounds[r + 1] - bounds[r], F), device=cuda) for r in range(3)]
            for r in range(3):
                with torch.cuda.stream(streams[r]):
                    ld.ranks[r].gcn_run(Y[r], 0, F)
            torch.cuda.synchronize()
            for r in range(3):
                ld.ranks[r].check()
            got = torch.cat(Y).cpu().numpy().astype(np.float64)
            assert np.all(np.abs(got - want) <= 1e-5 * scale + 1e-30), F
    finally:
        ld.close()


def test_state_errors_are_reported_not_crashed(gn, cuda):
    import ctypes as C

    L = gn.lib()
    bounds = (C.c_int64 * 3)(0, 8, 16)
    h = C.c_void_p()
    assert L.gnnagg_dist_create_rank(0, 2, bounds, 32, C.byref(h)) == 0
    blob = C.create_string_buffer(256)
    assert L.gnnagg_dist_export(h, blob) == -4                         # no graph yet
    assert L.gnnagg_dist_gcn_run(h, 0, None, 32, 0, None) != 0         # neither graph nor peers
    assert L.gnnagg_dist_x(h, 0) is None
    ptr = torch.zeros(9, dtype=torch.int32, device=cuda)
    assert L.gnnagg_dist_set_graph(h, C.c_void_p(ptr.data_ptr()), None, None, 0, 1, None) == 0
    assert L.gnnagg_dist_set_graph(h, C.c_void_p(ptr.data_ptr()), None, None, 0, 1, None) == -4   # once per handle
    Y = torch.empty((8, 32), device=cuda)
    assert L.gnnagg_dist_gcn_run(h, 0, C.c_void_p(Y.data_ptr()), 32, 0, None) == -4              # peers not connected
    assert b"connect" in L.gnnagg_last_error()
    assert L.gnnagg_dist_gcn_run(h, 0, C.c_void_p(Y.data_ptr()), 48, 0, None) != 0               # feat > feat_cap
    assert L.gnnagg_dist_create_rank(0, 17, bounds, 32, C.byref(C.c_void_p())) == -1              # world > 16
    assert L.gnnagg_dist_destroy(h) == 0


def test_two_layers_through_the_two_shard_buffers(gn, orc, cuda):
    """a 2-layer GCN across ranks without any copy between the layers: layer 1 writes H1 = (A X) W1 straight into the
    OTHER peer-visible shard buffer, layer 2 pushes its halo out of that buffer.  Exercises consecutive epochs, the
    write-after-read protection of the receive slots and both buffers."""
    rng = np.random.default_rng(23)
    bounds = [0, 900, 2100, 3000]
    n, F = bounds[-1], 64
    ptr, idx = synth.small_random_csr(n, 18.0, 29, hub=8000)
    val = (rng.random(len(idx)).astype(np.float32) + 0.1) / 20
    X = rng.standard_normal((n, F)).astype(np.float32)
    W1 = (rng.standard_normal((F, F)) / 8).astype(np.float32)
    W2 = (rng.standard_normal((F, F)) / 8).astype(np.float32)
    _, h1_64, _ = orc.gcn_layer_f64(ptr, idx, val, X, W1)
    ld = partition.LocalDist(bounds, F, devices=[0, 0, 0])
    streams = [torch.cuda.Stream(device=cuda) for _ in range(3)]
    try:
        for r in range(3):
            lp, li, lv = partition.local_block(ptr, idx, val, bounds, r)
            ld.set_graph(r, torch.from_numpy(lp).to(cuda), torch.from_numpy(li).to(cuda), torch.from_numpy(lv).to(cuda), 2)
        ld.connect()
        dW1, dW2 = torch.from_numpy(W1).to(cuda), torch.from_numpy(W2).to(cuda)
        H2 = [torch.empty((bounds[r + 1] - bounds[r], F), device=cuda) for r in range(3)]
        for rep in range(3):  # repeated: the buffers are rewritten while peers may still be behind
            for r in range(3):
                with torch.cuda.stream(streams[r]):
                    ld.ranks[r].x(0, F).copy_(torch.from_numpy(np.ascontiguousarray(X[bounds[r]:bounds[r + 1]])).to(cuda), non_blocking=True)
                    ld.ranks[r].gcn_layer(dW1, ld.ranks[r].x(1, F), buf=0)       # H1 lands in buffer 1 of every rank
                    ld.ranks[r].gcn_layer(dW2, H2[r], buf=1)                      # layer 2 reads buffer 1
        torch.cuda.synchronize()
        for r in range(3):
            ld.ranks[r].check()
        h1 = torch.cat([ld.ranks[r].x(1, F) for r in range(3)]).cpu().numpy()
        got = torch.cat(H2).cpu().numpy().astype(np.float64)
    finally:
        ld.close()
    # layer 2 of the oracle starts from the GPU's own H1 (fp32), so that only layer 2's error is gated
    _, h2_64, h2_scale = orc.gcn_layer_f64(ptr, idx, val, np.ascontiguousarray(h1), W2)
    assert np.all(np.abs(h1.astype(np.float64) - h1_64) <= 1e-4 * np.abs(h1_64).max())
    assert np.all(np.abs(got - h2_64) <= 1e-5 * h2_scale + 1e-30)
