#!/usr/bin/env python
"""Why overlapping the halo with the aggregation pays on the reddit-shaped layer and not on RMAT-26:
a ONE-GPU probe of the memory-system contention between the two halves of a rank's step, with NVLink taken out of
the picture.  It builds rank 0's row block of the N-rank workload, re-indexes it onto [shard | receive slots] as
gnnagg_dist_set_graph does, and times
    (a) the aggregation alone,
    (b) a local stand-in for the exchange alone: gather the rows this rank would PUSH out of its shard and write them
        to a buffer (the local half of halo_push_kernel) plus a streaming write of the bytes it would RECEIVE,
    (c) both at once, the stand-in on a high-priority stream.
If (c) is close to (a) + (b), the two halves fight for the same resource on this GPU and no schedule can hide one behind
the other.   python tools/contention_probe.py [rmat26|reddit] [N]   -> one JSON line"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gnn-computing_b200"))
import torch

import gnnagg
from gnnagg import synth


def main():
    shape = sys.argv[1] if len(sys.argv) > 1 else "rmat26"
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    dev = torch.device("cuda:0")
    if shape == "rmat26":
        n, m, F = (1 << 26) // N, (1 << 30) // N, 64
    else:
        n, m, F = 232965, 114615891, 128
    ptr, idx = synth.rmat_csr(n, m, seed=123, device=dev, src_num_v=n * N, dst_prefix=0)
    val = torch.rand(m, device=dev) + 0.5
    g = idx.long()
    remote = g >= n
    U = torch.unique(g[remote])
    idx_new = torch.where(remote, n + torch.searchsorted(U, g), g).to(torch.int32)
    R = U.numel()
    X = torch.randn((n + R, F), device=dev)                       # shard rows, then the receive slots
    Y = torch.empty((n, F), device=dev)
    agg = gnnagg.Aggregator(ptr, idx_new, val)
    # what this rank pushes: by symmetry as many rows as it receives, drawn from its own shard, ascending per receiver
    per = R // (N - 1)
    send = torch.cat([torch.sort(torch.randint(0, n, (per,), device=dev))[0] for _ in range(N - 1)])
    out = torch.empty((send.numel(), F), device=dev)
    incoming = torch.empty((R, F), device=dev)
    side = torch.cuda.Stream(device=dev, priority=-1)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def exchange():
        gnnagg.gather_rows(X, send, out)                            # local reads of the push
        incoming.fill_(1.0)                                         # the bytes that arrive over NVLink land in HBM

    def timeit(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            flush.fill_(1)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return sorted(ts)[len(ts) // 2]

    side_ev = [torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)]

    def both():
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            side_ev[0].record()
            exchange()
            side_ev[1].record()
        agg.gcn_run(X, Y)
        torch.cuda.current_stream().wait_stream(side)

    if len(sys.argv) > 3:
        agg.set_warp_edges(int(sys.argv[3]))     # 128: the small-graph variant, half the gathers in flight per warp
    t_agg = timeit(lambda: agg.gcn_run(X, Y))
    t_ex = timeit(exchange)
    t_both = timeit(both)
    torch.cuda.synchronize()
    t_ex_under = side_ev[0].elapsed_time(side_ev[1])   # how long the stand-in takes while the aggregation runs (last repetition)
    print(json.dumps({"probe": "local contention between aggregation and exchange stand-in", "shape": shape, "ranks": N, "rows": n,
                      "edges": m, "F": F, "recv_rows": R, "exchange_bytes_each_way": R * F * 4,
                      "agg_alone_ms": round(t_agg, 3), "exchange_standin_alone_ms": round(t_ex, 3), "both_ms": round(t_both, 3), "exchange_standin_while_aggregating_ms": round(t_ex_under, 3),
                      "warp_edges": int(sys.argv[3]) if len(sys.argv) > 3 else 512,
                      "sum_ms": round(t_agg + t_ex, 3), "overlap_gain_ms": round(t_agg + t_ex - t_both, 3)}))


if __name__ == "__main__":
    main()
