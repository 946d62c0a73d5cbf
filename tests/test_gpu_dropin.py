"""GPU tier: the drop-in boundary at run time.
  * tests/cpp/dropin_check.cu -- a caller written against the reference's C++ aggregator API, built
    against include/ + libgnnagg.so, validates every variant with the reference's own valid() utility;
  * the reference's UNMODIFIED drivers (Figure9/main.cu, Figure10/main_a.cu, main_b.cu), compiled from
    /root/reference against our include/ by tools/build_compat.sh, must run to completion on the
    reference's file formats (.config/.graph/.reorder, ../data relative to the CWD)."""
import os
import subprocess

import numpy as np
import pytest

from gnnagg import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "build", "compat")


@pytest.fixture(scope="module")
def dataset(gn, cuda, tmp_path_factory):
    top = tmp_path_factory.mktemp("refdata")
    (top / "data").mkdir()
    (top / "run").mkdir()
    ptr, idx = synth.rmat_csr(20000, 300000, seed=5)
    ptr, idx = ptr.numpy(), idx.numpy()
    d = str(top / "data") + "/"
    gn.write_graph("syn", ptr, idx, d)
    rows = np.random.default_rng(0).permutation(20000).astype(np.int32)
    gn.write_reorder(d + "syn.reorder_thres_0.2", rows)
    return str(top / "run")


def _run(binary, cwd, *flags):
    path = os.path.join(BIN, binary)
    if not os.path.exists(path):
        pytest.skip("%s not built (tools/build_compat.sh)" % binary)
    p = subprocess.run([path] + list(flags), cwd=cwd, capture_output=True, text=True, timeout=300)
    return p.returncode, p.stderr + p.stdout


@pytest.mark.parametrize("flags", [("--feature-len", "32", "--nei", "16"), ("--feature-len", "64", "--nei", "32"),
                                   ("--feature-len", "128", "--nei", "64", "--outfea", "32"),
                                   ("--feature-len=32", "--nei=32", "--reorder", "_thres_0.2")])
def test_dropin_check_driver(dataset, flags):
    rc, out = _run("dropin_check.out", dataset, "--dataset", "syn", *flags)
    assert rc == 0 and "DROPIN_CHECK ok" in out, out[-3000:]
    if "--reorder" in flags:
        assert "reorder:" in out


def test_reference_fig9_driver_runs_unmodified(dataset):
    rc, out = _run("Figure9_main.out", dataset, "--dataset", "syn", "--feature-len", "32")
    assert rc == 0, out[-2000:]
    assert "num_target" in out  # dbg(num_target) of Aggregator::schedule, scraped by the reference's scripts


def test_reference_fig10_drivers_run_unmodified(dataset):
    rc, out = _run("Figure10_main_a.out", dataset, "--dataset", "syn", "--feature-len", "32", "--nei", "32")
    assert rc == 0, out[-2000:]
    # Figure10/run.sh:17-19 greps these names and takes awk column 8
    for key in ("t_base", "t_adapter", "t_linear"):
        line = [l for l in out.splitlines() if key in l][0]
        assert float(line.split()[7]) > 0
    rc, out = _run("Figure10_main_b.out", dataset, "--dataset", "syn", "--feature-len", "32", "--outfea", "32", "--nei", "64")
    assert rc == 0, out[-2000:]
    for key in ("t_base", "t_linear"):
        line = [l for l in out.splitlines() if key in l][0]
        assert float(line.split()[7]) > 0
