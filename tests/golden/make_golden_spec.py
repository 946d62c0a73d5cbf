"""Generates tests/golden/spec_vectors.json: known-answer vectors for the parts of the product whose specification is
THIS repository's (the reference has no usable counterpart): the fixed-fanout sampler's position function
(orc_sample_pos), a sampled sub-graph, the stable CSR transpose and the GAT backward of a 6-vertex graph.
They freeze the specification: tests/test_spec_golden.py checks the oracle against them on CPU, the GPU tier checks
the CUDA path against the oracle.      python tests/golden/make_golden_spec.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    out = {}
    out["sample_pos"] = [{"seed": s, "v": v, "j": j, "deg": d, "fanout": k, "pos": O.sample_pos(s, v, j, d, k)}
                         for (s, v, j, d, k) in [(123, 0, 0, 17, 4), (123, 0, 3, 17, 4), (123, 7, 2, 1000, 16), (5, 7, 2, 1000, 16),
                                                 (0, 2 ** 31 - 1, 15, 2 ** 30, 16), (77, 41, 0, 2, 1), (9, 3, 9, 33, 10)]]
    ptr = np.array([0, 3, 3, 8, 9, 15, 15], np.int32)
    idx = np.array([1, 2, 3, 0, 1, 2, 3, 5, 2, 0, 1, 2, 3, 4, 5], np.int32)
    active = np.array([0, 0, 1, 0, 0, 0], np.int32)
    act, vs, sp, si = O.sample_subgraph(ptr, idx, active, 2, 2, seed=123)
    out["sample_subgraph"] = {"ptr": ptr.tolist(), "idx": idx.tolist(), "active": active.tolist(), "fanout": 2, "layers": 2,
                              "seed": 123, "active_after": act.tolist(), "vertexset": vs.tolist(), "sub_ptr": sp.tolist(),
                              "sub_idx": si.tolist()}
    t_ptr, t_idx, t_perm = O.transpose_csr(ptr, idx, 6)
    out["transpose"] = {"t_ptr": t_ptr.tolist(), "t_idx": t_idx.tolist(), "t_perm": t_perm.tolist()}
    rng = np.random.default_rng(20261017)
    X = rng.standard_normal((6, 4)).astype(np.float32)
    dY = rng.standard_normal((6, 4)).astype(np.float32)
    att = rng.standard_normal((6, 2)).astype(np.float32)
    dX, dA, _, _ = O.gat_backward_f64(ptr, idx, att, X, dY, 0.2)
    out["gat_backward"] = {"X": X.tolist(), "dY": dY.tolist(), "att": att.tolist(), "slope": 0.2, "dX": dX.astype(np.float64).tolist(),
                           "datt": dA.astype(np.float64).tolist()}
    with open(os.path.join(HERE, "spec_vectors.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote spec_vectors.json")


if __name__ == "__main__":
    main()
