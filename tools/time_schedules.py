#!/usr/bin/env python
"""Host-side timing of rows a3-a5 (the three schedules): the product's host builders (gnnagg_schedule_build) beside the
reference's own host functions compiled into oracle/_ref/libref.so, same graph, outputs compared byte for byte.  CPU only.
usage: tools/time_schedules.py [num_v] [num_e]   -> JSON lines"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gnn-computing_b200"))
import numpy as np

import gnnagg
import oracle
from gnnagg import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 20_000_000
ptr, idx = synth.rmat_csr(n, m, seed=123)
ptr, idx = ptr.numpy(), idx.numpy()
val = np.random.default_rng(0).standard_normal(m).astype(np.float32)
for kind, name, kw in ((1, "neighbor_grouping(32)", dict(neighbor_num=32)), (0, "locality(8)", dict(par_num=8)),
                       (2, "locality_neighbor_grouping(8,32)", dict(par_num=8, neighbor_num=32))):
    t0 = time.time()
    ours = gnnagg.schedule_build(kind, ptr, idx, None if kind == 1 else val, **kw)
    t_ours = time.time() - t0
    out = {"num_v": n, "num_e": m, "schedule": name, "ours_host_s": round(t_ours, 3), "groups": int(len(ours[2]))}
    if oracle.ref_available():
        t0 = time.time()
        ref = oracle.ref_schedule(kind, ptr, idx, None if kind == 1 else val, total_num_v=n, **kw)
        out["ref_host_s"] = round(time.time() - t0, 3)
        out["identical"] = all((a is None and b is None) or np.array_equal(a, b) for a, b in zip(ours, ref))
    print(json.dumps(out))
