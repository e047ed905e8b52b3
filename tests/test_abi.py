"""The C-ABI library loads and exports every symbol include/protpore_b200.h declares; without a GPU the
compute path fails loudly (there is no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import util
from protpore_b200 import _engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "protpore_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_engine.LIB_PATH)
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), n
    assert lib.pp_abi_version() == 1


def test_python_binding_covers_the_header():
    assert sorted(n for n, _, _ in _engine.ABI) == _declared()
    _engine.load_library()


def test_no_gpu_fails_loudly():
    if _engine.device_count() > 0:
        pytest.skip("a CUDA device is present")
    model = util.build_model("toy")
    with pytest.raises(_engine.EngineError) as info:
        model.viterbi(np.array([0.1, 0.2]))
    assert info.value.code == _engine.PP_ERR_CUDA
    assert "no CPU fallback" in str(info.value)
    with pytest.raises(_engine.EngineError):
        model.log_probability_batch([np.array([0.1])])


def test_product_never_imports_the_oracle():
    """Nothing under protpore_b200/ may reference oracle/ (the oracle is test infrastructure)."""
    pkg = os.path.join(ROOT, "protpore_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(base, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "hmm_oracle" not in text and "oracle/_ref" not in text, f
