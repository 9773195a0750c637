"""CPU tests of the drop-in boundary: the C-ABI library builds, loads and exports every declared symbol."""
import ctypes
import os
import re

import pytest
import torch

import tamtr_b200
from tamtr_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "tamtr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tamtr_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    _lib.build()
    assert os.path.exists(_lib.LIB_PATH)
    handle = ctypes.CDLL(_lib.LIB_PATH)       # loads without a GPU (libcudart only)
    declared = _declared_symbols()
    assert len(declared) >= 6
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in include/tamtr_b200.h but not exported"
    assert sorted(_lib.exported_symbols()) == declared, "ctypes binding and header are out of sync"
    assert handle.tamtr_abi_version() == 2


def test_cpu_tensors_are_rejected_with_the_message_the_reference_expects():
    # ultralytics/nn/tasks.py:256-264 moves the model to CUDA iff the error mentions one of these strings
    v = torch.zeros(1, 4, 1, 8)
    loc = torch.zeros(1, 1, 1, 1, 1, 2)
    a = torch.ones(1, 1, 1, 1, 1)
    with pytest.raises(RuntimeError) as e:
        tamtr_b200.ms_deform_attn(v, [[2, 2]], loc, a)
    assert "Not implemented on the CPU" in str(e.value) and "is_cuda" in str(e.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "tamtr_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src.replace(
                    "never imports `oracle/`", ""), f"{f} references the oracle"


def test_argument_validation_without_gpu():
    lib = _lib.lib()
    sh = (ctypes.c_int32 * 2)(2, 2)
    # null pointers -> TAMTR_E_BADARG before any CUDA call
    rc = lib.tamtr_msda_forward(None, None, None, None, 0, 1, 4, 1, 8, 1, 1, 1, sh, 0, None)
    assert rc == -1 and b"null" in lib.tamtr_last_error()
    # unsupported head dim
    rc = lib.tamtr_msda_forward(1, 1, 1, 1, 0, 1, 4, 1, 7, 1, 1, 1, sh, 0, None)
    assert rc == -2
    # level shapes do not add up to Lv
    rc = lib.tamtr_msda_forward(1, 1, 1, 1, 0, 1, 5, 1, 8, 1, 1, 1, sh, 0, None)
    assert rc == -1
    # token stride smaller than H*Dh
    rc = lib.tamtr_msda_forward(1, 1, 1, 1, 0, 1, 4, 1, 8, 1, 1, 1, sh, 4, None)
    assert rc == -1
