"""CPU: the C-ABI library loads and exports every symbol include/bq.h declares; argument validation that
needs no GPU behaves as documented.  (No compute calls here.)"""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from llm_mixed_q_b200 import _lib as L


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "bq.h")).read()
    return sorted(set(re.findall(r"BQ_API\s+[\w\s\*]+?\b(bq_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = L.load()
    names = declared_symbols()
    assert len(names) >= 10
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/bq.h but not exported"


def test_abi_version_and_strerror():
    lib = L.load()
    src = open(os.path.join(ROOT, "include", "bq.h")).read()
    assert lib.bq_abi_version() == int(re.search(r"#define BQ_ABI_VERSION (\d+)", src).group(1))
    assert lib.bq_strerror(0) == b"ok"
    for code in range(1, 7):
        assert lib.bq_strerror(code) not in (b"ok", b"unknown status")
    assert lib.bq_strerror(99) == b"unknown status"


def test_struct_layouts_match_header():
    assert ctypes.sizeof(L.BqFormat) == 32
    assert ctypes.sizeof(L.BqTensor3) == 48


def test_bad_arguments_are_rejected_without_a_gpu():
    lib = L.load()
    t = L.BqTensor3(1, 4, 64, 256, 64, 1)
    # mantissa bits < 0  -> BQ_ERR_BAD_FORMAT
    f = L.BqFormat(L.KIND["block_minifloat"], 3, 4, 0, 8, 1, 16, 0)
    assert lib.bq_quantize(ctypes.byref(f), ctypes.byref(t), 16, 16, 0, 0, None, 0, None) == 3
    with pytest.raises(ValueError):
        L.check(3, "x")
    # null data pointers -> BQ_ERR_BAD_ARG
    f = L.BqFormat(L.KIND["block_fp"], 6, 8, 127, 0, 1, 16, 0)
    assert lib.bq_quantize(ctypes.byref(f), ctypes.byref(t), None, None, 0, 0, None, 0, None) == 1
    # negative size
    tn = L.BqTensor3(1, -4, 64, 256, 64, 1)
    assert lib.bq_quantize(ctypes.byref(f), ctypes.byref(tn), 16, 16, 0, 0, None, 0, None) == 1
    # workspace query is pure host arithmetic
    assert lib.bq_quantize_workspace_bytes(ctypes.byref(f), ctypes.byref(t)) >= 16
    fl = L.BqFormat(L.KIND["block_log"], 8, 0, 0, 8, 1, 16, 0)
    big = L.BqTensor3(1, 1024, 4096, 1024 * 4096, 4096, 1)
    assert lib.bq_quantize_workspace_bytes(ctypes.byref(fl), ctypes.byref(big)) >= 1024 * 4096 // 32
    # GEMM: misaligned leading dimension -> BQ_ERR_BAD_ARG ; K == 0 -> unsupported
    assert lib.bq_gemm_bf16_tn(256, 256, 256, None, 1, 8, 8, 12, 12, 12, 8, 0, 0, 0, None) == 1
    assert lib.bq_gemm_bf16_tn(256, 256, 256, None, 1, 8, 8, 0, 8, 8, 8, 0, 0, 0, None) == 2
    # linear with a format that is not bf16 exact (block_fp width 12)
    f12 = L.BqFormat(L.KIND["block_fp"], 12, 8, 127, 0, 1, 16, 0)
    assert lib.bq_linear(ctypes.byref(f12), 256, 8, 64, 64, 256, 8, None, 256, 8, 256, 1 << 20, None) == 6
    with pytest.raises(NotImplementedError):
        L.check(6, "x")


def test_no_cpu_fallback():
    import torch

    from llm_mixed_q_b200.models.quantize.quantizers import block_fp_quantizer

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        block_fp_quantizer(torch.randn(4, 32), 6, 8, 127, [1, 16], True)
