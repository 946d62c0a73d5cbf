"""CPU tier: the C-ABI shared library loads without a GPU and exports every symbol declared in
include/gnnagg.h; device entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "gnnagg.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gnnagg_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(gn):
    names = _declared()
    assert len(names) >= 45
    L = C.CDLL(gn.LIB_PATH)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    # the python binding covers the whole header as well
    assert sorted(gn.exported_symbols()) == names


def test_header_is_plain_c(tmp_path):
    """the header must compile as C (no C++ / torch types in the signatures)"""
    src = tmp_path / "t.c"
    src.write_text('#include "gnnagg.h"\nint main(void){return gnnagg_version()==0;}\n')
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)])


def test_version_and_error_string(gn):
    assert gn.lib().gnnagg_version() >= 100
    with pytest.raises(gn.GnnaggError) as e:
        gn.schedule_build(1, np.array([0, 1], np.int32), np.array([0], np.int32), neighbor_num=-1)
    assert "neighbor_num" in str(e.value)


def test_device_entry_points_fail_loudly_without_gpu(gn):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(gn.GnnaggError):
        gn.device_info()
    h = C.c_void_p()
    fake = C.c_void_p(0x1000)
    rc = gn.lib().gnnagg_create(fake, fake, None, None, 4, 9, C.byref(h))
    assert rc == -2 and b"cuda" in gn.lib().gnnagg_last_error().lower()  # GNNAGG_ERR_CUDA, nothing computed
    with pytest.raises(gn.GnnaggError):
        gn.Aggregator(torch.zeros(2, dtype=torch.int32), torch.zeros(1, dtype=torch.int32))


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under the product package may reference it"""
    pkg = os.path.join(ROOT, "gnn-computing_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in text and "liboracle" not in text and "libref" not in text, f
