"""The C-ABI library loads and exports exactly what include/sr2.h declares
(no compute calls: this runs without a GPU)."""
import ctypes
import os
import re

import pytest

from superrec2_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "sr2.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(sr2_[a-z_0-9]+)\s*\(", text))


def test_header_symbols_are_exported_and_bound():
    declared = _declared()
    assert declared, "no prototypes found in include/sr2.h"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in sr2.h but not exported by libsr2.so"
    assert declared == set(_lib.SIGNATURES), "ctypes signatures out of sync with include/sr2.h"


def test_no_gpu_fails_loudly():
    """Without a usable GPU the product refuses to run (no CPU fallback)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.Sr2Error) as err:
        _lib.Context(0)
    assert "no CPU path" in str(err.value) or "CUDA" in str(err.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "superrec2_b200")
    for folder, _, files in os.walk(pkg):
        for name in files:
            if name.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
                text = open(os.path.join(folder, name)).read()
                assert "oracle" not in text.lower().replace("lowestcommonancestor oracle", "") \
                    .replace("lca oracle", ""), f"{name} mentions the oracle"
