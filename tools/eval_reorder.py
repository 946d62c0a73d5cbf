#!/usr/bin/env python
"""C4 evaluation: locality reorder (gnnagg_lsh_reorder = cluster2.py semantics) + neighbour-grouped /
locality schedules on the products-shaped graph, F = 256.  Times the aggregation on the original and on
the reordered graph (un-scheduled, NG=32, locality+NG(8,32)); run it under
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,gpu__time_duration.sum \
      -k regex:agg_kernel --csv --log-file gpurun_out/r1_reorder_ncu.csv python tools/eval_reorder.py --reps 1
to get DRAM bytes / L2 hit rate per variant (launch order = order of the `variants` list)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gnn-computing_b200"))

import numpy as np
import torch

import gnnagg
from gnnagg import synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="products")
    ap.add_argument("--feat", type=int, default=256)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r1_reorder_eval.json"))
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    if args.shape == "planted":  # community-structured graph with scrambled ids (what the reorder targets)
        n = 1200000
        ptr, idx = synth.planted_community_csr(n, seed=123, device=dev)
        m = int(idx.numel())
    else:
        n, m = synth.shape_of(args.shape)
        ptr, idx = synth.rmat_csr(n, m, seed=123, device=dev)
    val = synth.gcn_norm_val(ptr, idx)
    hp, hi, hv = ptr.cpu().numpy(), idx.cpu().numpy(), val.cpu().numpy()
    t0 = time.time()
    rows = gnnagg.lsh_reorder(hp, hi)
    t_reorder = time.time() - t0
    rev = np.empty(n, np.int32)
    rev[rows] = np.arange(n, dtype=np.int32)
    t0 = time.time()
    rp, ri = gnnagg.reorder_csr(hp, hi, rows, rev)
    t_apply = time.time() - t0
    # edge values follow their edges: new row i = old row rows[i], same within-row order (src/data.cu:19-24)
    starts = hp[:-1][rows].astype(np.int64)
    lens = np.diff(hp)[rows].astype(np.int64)
    offs = np.repeat(starts - np.concatenate([[0], np.cumsum(lens)[:-1]]), lens) + np.arange(len(hi), dtype=np.int64)
    rv = hv[offs]
    X = torch.randn((n, args.feat), device=dev, generator=torch.Generator(device=dev).manual_seed(123))
    Y = torch.empty((n, args.feat), device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    graphs = {"original": (ptr, idx, val), "reordered": tuple(torch.from_numpy(a).to(dev) for a in (rp, ri, rv))}
    # clustering quality: how many consecutive rows share their most frequent source
    out = {"shape": args.shape, "n": n, "m": m, "F": args.feat, "reorder_s": round(t_reorder, 2), "apply_s": round(t_apply, 2),
           "gather_model_bytes": 4 * (n + 1) + 8 * m + 4 * m * args.feat + 4 * n * args.feat, "variants": []}

    def timeit(fn):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.reps):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.median(ts))

    for gname, (p, i, v) in graphs.items():
        agg = gnnagg.Aggregator(p, i, v)
        for sname, kind, params in (("unscheduled", None, None), ("neighbor_grouping_32", 1, [32]),
                                    ("locality_neighbor_grouping_8_32", 2, [8, 32])):
            t_sched = 0.0
            if kind is not None:
                t0 = time.time()
                agg.schedule(kind, params)
                t_sched = time.time() - t0
            ms = timeit(lambda: agg.gcn_run(X, Y, scheduled=kind is not None))
            out["variants"].append({"graph": gname, "schedule": sname, "ms": round(ms, 4), "schedule_s": round(t_sched, 2),
                                    "GBps": round(out["gather_model_bytes"] / ms / 1e6, 1)})
            print(out["variants"][-1], flush=True)
        agg.close()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
