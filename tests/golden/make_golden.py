"""Generates tests/golden/*.json by running the REFERENCE's own host functions.

Run in the build container (needs /root/reference, compiled by `make -C oracle` into
oracle/_ref/libref.so):      python tests/golden/make_golden.py
The produced vectors pin the CPU oracle (oracle/oracle.c) and the product's host preprocessing
(gnnagg_schedule_build, gnnagg_reorder_csr, gnnagg_graph_load); they are small on purpose.
Functions exercised: neighbor_grouping_schedule / locality_schedule / localityNeighborGrouping
(include/graph_schedule.h:17-243), reorderCSR and load_graph (src/data.cu:4-139).
"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def rand_csr(rng, n, avg, empty_frac, hub=0):
    deg = rng.poisson(avg, n)
    deg[rng.random(n) < empty_frac] = 0
    if hub:
        deg[rng.integers(n)] = hub
    ptr = np.zeros(n + 1, np.int32)
    ptr[1:] = np.cumsum(deg)
    idx = rng.integers(0, n, int(ptr[-1])).astype(np.int32)
    return ptr, idx


def cases():
    # the hand-checked case of SURVEY.md section 4
    yield "survey4", np.array([0, 3, 3, 8, 9], np.int32), np.array([1, 2, 3, 0, 1, 2, 3, 0, 2], np.int32)
    rng = np.random.default_rng(20261017)
    yield "single_empty_row", np.array([0, 0], np.int32), np.zeros(0, np.int32)
    yield "all_empty", np.zeros(6, np.int32), np.zeros(0, np.int32)
    yield "one_full_row", np.array([0, 7], np.int32), np.array([0, 0, 0, 0, 0, 0, 0], np.int32)
    for k, (n, avg, ef, hub) in enumerate([(7, 3, 0.3, 0), (33, 5, 0.2, 40), (48, 12, 0.0, 0), (50, 2, 0.5, 97),
                                           (90, 6, 0.1, 200)]):
        ptr, idx = rand_csr(rng, n, avg, ef, hub)
        yield "rand%d" % k, ptr, idx


def main():
    if not O.ref_available():
        raise SystemExit("oracle/_ref/libref.so missing: run `make -C oracle` in the build container")
    out = {"generator": "tests/golden/make_golden.py", "source": "reference host functions via oracle/_ref/libref.so",
           "cases": []}
    rng = np.random.default_rng(7)
    for name, ptr, idx in cases():
        n, m = len(ptr) - 1, len(idx)
        val = rng.integers(1, 64, m).astype(np.float32) / 8  # exact in fp32, short in JSON
        case = {"name": name, "ptr": ptr.tolist(), "idx": idx.tolist(), "val": [float(v) for v in val], "schedules": []}
        for ng in (1, 2, 16):
            p, i, t, _ = O.ref_schedule(1, ptr, idx, neighbor_num=ng)
            case["schedules"].append({"kind": 1, "neighbor_num": ng, "ptr": p.tolist(), "idx": i.tolist(),
                                      "target": t.tolist()})
        for par in (1, 2, 3):
            for total in sorted({n, max(n - 1, 1)}):
                p, i, t, v = O.ref_schedule(0, ptr, idx, val, par_num=par, total_num_v=total)
                case["schedules"].append({"kind": 0, "par_num": par, "total_num_v": total, "ptr": p.tolist(),
                                          "idx": i.tolist(), "target": t.tolist(), "val": [float(x) for x in v]})
                for ng in (2, 16):
                    p, i, t, v = O.ref_schedule(2, ptr, idx, val, par_num=par, neighbor_num=ng, total_num_v=total)
                    case["schedules"].append({"kind": 2, "par_num": par, "neighbor_num": ng, "total_num_v": total,
                                              "ptr": p.tolist(), "idx": i.tolist(), "target": t.tolist(),
                                              "val": [float(x) for x in v]})
        rows = rng.permutation(n).astype(np.int32)
        rev = np.empty(n, np.int32)
        rev[rows] = np.arange(n, dtype=np.int32)
        np_, ni = O.ref_reorder_csr(ptr, idx, rows, rev)
        case["reorder"] = {"rows": rows.tolist(), "reverse_rows": rev.tolist(), "newptr": np_.tolist(),
                           "newidx": ni.tolist()}
        out["cases"].append(case)

    # load_graph: text parse + cache files + reorder, run from a scratch tree laid out as the
    # reference expects (CWD = Figure*/, files under ../data/)
    ptr, idx = np.array([0, 3, 3, 8, 9], np.int32), np.array([1, 2, 3, 0, 1, 2, 3, 0, 2], np.int32)
    with tempfile.TemporaryDirectory() as td:
        os.makedirs(os.path.join(td, "data"))
        os.makedirs(os.path.join(td, "run"))
        with open(os.path.join(td, "data", "tiny.config"), "w") as f:
            f.write("4 9")
        with open(os.path.join(td, "data", "tiny.graph"), "w") as f:
            f.write(" ".join(map(str, ptr)) + "\n" + " ".join(map(str, idx)) + "\n")
        with open(os.path.join(td, "data", "tiny.reorder_t"), "w") as f:
            f.write("2 0 3 1 ")
        cwd = os.getcwd()
        os.chdir(os.path.join(td, "run"))
        try:
            lp, li, rows, rev = O.ref_load_graph("tiny", "_t")
            dumps = {k: list(np.fromfile(os.path.join(td, "data", "tiny.graph." + k), np.int32).tolist())
                     for k in ("ptrdump", "edgedump")}
            lp2, li2, rows2, _ = O.ref_load_graph("tiny", "")  # now served from the caches, no reorder
        finally:
            os.chdir(cwd)
    out["load_graph"] = {"config": "4 9", "graph_ptr": ptr.tolist(), "graph_idx": idx.tolist(), "reorder_text": "2 0 3 1 ",
                         "reordered_ptr": lp.tolist(), "reordered_idx": li.tolist(), "rows": rows.tolist(),
                         "reverse_rows": rev.tolist(), "ptrdump": dumps["ptrdump"], "edgedump": dumps["edgedump"],
                         "cached_ptr": lp2.tolist(), "cached_idx": li2.tolist(), "cached_reordered": rows2 is not None}
    with open(os.path.join(HERE, "host_prep.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", os.path.join(HERE, "host_prep.json"), os.path.getsize(os.path.join(HERE, "host_prep.json")), "bytes")


if __name__ == "__main__":
    main()
