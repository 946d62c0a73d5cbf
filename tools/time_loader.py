#!/usr/bin/env python
"""Host-side timing of row a16 (load_graph): parse of the reference's text format and the raw-dump fast path, ours
(libgnnagg.so) beside the reference's own load_graph compiled into oracle/_ref/libref.so.  CPU only.
usage: tools/time_loader.py [num_v] [num_e] [workdir]   -> one JSON line"""
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
work = sys.argv[3] if len(sys.argv) > 3 else "/tmp/gnnagg_loader"
os.makedirs(work + "/data", exist_ok=True)
os.makedirs(work + "/run", exist_ok=True)
ptr, idx = synth.rmat_csr(n, m, seed=123)
ptr, idx = ptr.numpy(), idx.numpy()
d = work + "/data/"


def clean():
    for suffix in (".graph.ptrdump", ".graph.edgedump"):
        if os.path.exists(d + "syn" + suffix):
            os.remove(d + "syn" + suffix)


t0 = time.time()
gnnagg.write_graph("syn", ptr, idx, d)
out = {"num_v": n, "num_e": m, "text_bytes": os.path.getsize(d + "syn.graph"), "ours_write_text_s": round(time.time() - t0, 3)}
clean()
os.chdir(work + "/run")
t0 = time.time()
p2, i2, _, _ = gnnagg.load_graph("syn", "../data/")
out["ours_parse_text_s"] = round(time.time() - t0, 3)
assert np.array_equal(p2, ptr) and np.array_equal(i2, idx)
t0 = time.time()
p3, i3, _, _ = gnnagg.load_graph("syn", "../data/")
out["ours_raw_dump_s"] = round(time.time() - t0, 3)
assert np.array_equal(p3, ptr) and np.array_equal(i3, idx)
if oracle.ref_available():
    clean()
    t0 = time.time()
    rp, ri, _, _ = oracle.ref_load_graph("syn")
    out["ref_parse_text_s"] = round(time.time() - t0, 3)
    assert np.array_equal(rp, ptr) and np.array_equal(ri, idx)
    t0 = time.time()
    rp, ri, _, _ = oracle.ref_load_graph("syn")
    out["ref_raw_dump_s"] = round(time.time() - t0, 3)
print(json.dumps(out))
