"""CPU tier: known-answer vectors for the repository's own specifications (fixed-fanout sampler, stable transpose, GAT
backward) -- tests/golden/spec_vectors.json, written by tests/golden/make_golden_spec.py.  A change of a hash constant,
of the stratum arithmetic or of the derivative shows up here before it silently moves both the oracle and the CUDA path."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def spec():
    with open(os.path.join(HERE, "golden", "spec_vectors.json")) as f:
        return json.load(f)


def test_sample_pos_vectors(orc, spec):
    for c in spec["sample_pos"]:
        assert orc.sample_pos(c["seed"], c["v"], c["j"], c["deg"], c["fanout"]) == c["pos"], c


def test_sample_subgraph_vector(orc, spec):
    c = spec["sample_subgraph"]
    got = orc.sample_subgraph(np.array(c["ptr"], np.int32), np.array(c["idx"], np.int32), np.array(c["active"], np.int32),
                              c["fanout"], c["layers"], seed=c["seed"])
    for g, key in zip(got, ("active_after", "vertexset", "sub_ptr", "sub_idx")):
        assert g.tolist() == c[key], key


def test_transpose_and_gat_backward_vectors(orc, spec):
    s = spec["sample_subgraph"]
    ptr, idx = np.array(s["ptr"], np.int32), np.array(s["idx"], np.int32)
    t = spec["transpose"]
    got = orc.transpose_csr(ptr, idx, 6)
    assert [g.tolist() for g in got] == [t["t_ptr"], t["t_idx"], t["t_perm"]]
    b = spec["gat_backward"]
    dX, dA, _, _ = orc.gat_backward_f64(ptr, idx, np.array(b["att"], np.float32), np.array(b["X"], np.float32),
                                        np.array(b["dY"], np.float32), b["slope"])
    np.testing.assert_allclose(dX, np.array(b["dX"]), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(dA, np.array(b["datt"]), rtol=1e-6, atol=1e-7)
