"""pytest configuration: `gpu` marker, import paths, shared fixtures.

CPU tier  (`-m "not gpu"`): oracle vs golden vectors / compiled reference, host preprocessing of
                            libgnnagg.so vs oracle, C-ABI symbol coverage, gloo multi-process logic.
GPU tier  (`-m gpu`)      : the parity tests proper -- CUDA path through the C ABI vs the oracle.
"""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gnn-computing_b200"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    import oracle

    oracle.lib()  # builds oracle/_build/liboracle.so on first use
    return oracle


@pytest.fixture(scope="session")
def gn():
    import gnnagg

    if not os.path.exists(gnnagg.LIB_PATH):
        gnnagg.build()
    gnnagg.lib()
    return gnnagg


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "host_prep.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible (there is no CPU fallback)")
    return torch.device("cuda:0")


def rel_gate(y, y64, scale, tol=1e-5):
    """parity gate of SURVEY.md 8(d): |y - y64| <= tol * sum|terms| element-wise (+ 1 ulp-ish floor)"""
    y = np.asarray(y, np.float64)
    err = np.abs(y - np.asarray(y64, np.float64))
    bound = tol * np.asarray(scale, np.float64) + 1e-30
    bad = err > bound
    worst = float((err / bound).max()) if err.size else 0.0
    return int(bad.sum()), worst
