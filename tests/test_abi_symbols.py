"""The C-ABI library loads on a CPU-only box and exports every symbol include/mevi_b200.h declares."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "mevi_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mevi_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_bound_and_exported():
    from mevi_b200 import _lib

    declared = _declared()
    assert len(declared) >= 15
    assert sorted(_lib.SYMBOLS) == declared, "ctypes table and header disagree"
    lib = _lib.load_library()  # binds every symbol; AttributeError if one is missing
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.mevi_abi_version() == 1


def test_no_torch_types_in_abi():
    text = open(os.path.join(ROOT, "include", "mevi_b200.h")).read()
    code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)  # declarations only, comments stripped
    assert "torch" not in code.lower() and "at::" not in code and "Tensor" not in code
    assert re.findall(r"#include\s*<([^>]+)>", code) == ["stdint.h"]


def test_missing_library_fails_loudly(tmp_path):
    env = dict(os.environ, MEVI_B200_LIB=str(tmp_path / "nope.so"), PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-c", "import mevi_b200; mevi_b200.load_library()"], env=env,
                       capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import mevi_b200

    with pytest.raises(mevi_b200.MeviError):
        mevi_b200.get_context()
    from mevi_b200.pq import ProductQuantization
    import numpy as np

    pq = ProductQuantization("rq", 2, 3, "l2", 16, "kmeans")
    with pytest.raises(mevi_b200.MeviError):
        pq.get_document_cluster(np.zeros((4, 16), np.float32), 0, 1)


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "mevi_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"
