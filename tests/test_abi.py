"""CPU: the C-ABI library loads, exports every symbol include/l32_ffn.h declares, and validates arguments
before touching the device (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

from llama32_b200 import _lib


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _lib.lib()


def test_header_symbols_exported(lib):
    names = _lib.header_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/l32_ffn.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes signature table out of sync with the header"


def test_header_arg_counts_match_ctypes_table():
    text = open(_lib.HEADER_PATH).read()
    for name, (_, argtypes) in _lib.SIGNATURES.items():
        m = re.search(r"L32_API[\w\s\*]*?\b" + name + r"\s*\(([^;]*?)\)\s*;", text, re.S)
        assert m, name
        args = m.group(1).strip()
        n = 0 if args in ("", "void") else len(args.split(","))
        assert n == len(argtypes), f"{name}: header has {n} parameters, ctypes table {len(argtypes)}"


def test_header_cites_reference_interfaces():
    text = open(_lib.HEADER_PATH).read()
    for cite in ("Tools/rmsnorm/rmsnorm.cu:7-32", "Tools/rmsnorm/rmsnorm.cu:35-61", "Tools/swiglu/swiglu.cu:277-316",
                 "Tools/swiglu/swiglu.cuh:18-25", "Tools/swiglu/swiglu.cu:319-364", "Model/model.py:214-217"):
        assert cite in text, cite


def test_version_and_error_strings(lib):
    assert lib.l32_abi_version() == 3
    assert lib.l32_error_string(0) == b"success"
    for code in (-1, -2, -3, -4, -5, -6, -7):
        assert len(lib.l32_error_string(code)) > 8


def test_argument_validation_without_device(lib):
    null = ctypes.c_void_p(0)
    one = ctypes.c_void_p(256)   # never dereferenced: validation fails first
    # bad dtype
    assert lib.l32_add_rmsnorm_forward(one, null, one, one, null, null, 4, 64, 1e-5, 7, null) == -1
    assert lib.l32_swiglu_forward(one, one, one, null, null, one, null, null, 4, 64, 128, 9, null) == -1
    # bad shapes (hidden / inter not a multiple of 8, negative tokens)
    assert lib.l32_swiglu_forward(one, one, one, null, null, one, null, null, 4, 60, 128, 0, null) == -2
    assert lib.l32_swiglu_forward(one, one, one, null, null, one, null, null, -1, 64, 128, 0, null) == -2
    assert lib.l32_linear_forward(one, one, null, one, 4, 64, 100, 0, null) == -2
    # null pointers
    assert lib.l32_swiglu_forward(null, one, one, null, null, one, null, null, 4, 64, 128, 0, null) == -4
    assert lib.l32_swiglu_forward(one, one, one, null, null, one, one, null, 4, 64, 128, 0, null) == -4  # one cache only
    assert lib.l32_add_rmsnorm_forward(null, null, one, one, null, null, 4, 64, 1e-5, 0, null) == -4
    # workspace checks
    assert lib.l32_rmsnorm_backward(one, one, one, one, one, null, null, 0, 4, 64, 0, null) == -6
    assert lib.l32_ffn_backward(one, one, one, one, one, one, one, one, null, null, null, null, 0, 4, 64, 128, 0, null) == -6
    # empty inputs are a no-op success
    assert lib.l32_add_rmsnorm_forward(null, null, null, null, null, null, 0, 64, 1e-5, 0, null) == 0
    assert lib.l32_swiglu_forward(null, null, null, null, null, null, null, null, 0, 64, 128, 0, null) == 0
    assert lib.l32_linear_forward(null, null, null, null, 0, 64, 128, 0, null) == 0


def test_workspace_sizes(lib):
    assert lib.l32_swiglu_backward_workspace_bytes(8192, 14336) >= 2 * 8192 * 14336 * 2
    assert lib.l32_ffn_backward_workspace_bytes(8192, 14336) >= 3 * 8192 * 14336 * 2
    assert lib.l32_rmsnorm_backward_workspace_bytes(8192, 4096) >= 4096 * 4
    assert lib.l32_swiglu_backward_workspace_bytes(0, 128) == 0


def test_extension_entry_points_importable():
    import rmsnorm
    import swiglu_fused
    assert callable(rmsnorm.forward) and callable(rmsnorm.backward)
    assert callable(swiglu_fused.forward) and callable(swiglu_fused.backward) and callable(swiglu_fused.forward_down)


def test_cuda_path_fails_loudly_on_cpu_tensors():
    import torch
    from llama32_b200 import ops
    x = torch.randn(4, 64, dtype=torch.bfloat16)
    w = torch.ones(64, dtype=torch.bfloat16)
    with pytest.raises(_lib.L32Error):
        ops.add_rmsnorm_forward(x, w, None, 1e-5)
    with pytest.raises(_lib.L32Error):
        ops.swiglu_forward(x, torch.randn(128, 64, dtype=torch.bfloat16), torch.randn(128, 64, dtype=torch.bfloat16))
