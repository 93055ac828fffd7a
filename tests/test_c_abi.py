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
    # entry points added later: argument checks run before any CUDA call
    f6 = L.BqFormat(L.KIND["block_fp"], 6, 8, 127, 0, 1, 16, 0)
    fl = L.BqFormat(L.KIND["block_log"], 4, 0, 0, 8, 1, 16, 0)
    q = ctypes.c_void_p
    # rope + quantise: null pointers, S % 16 != 0, head_dim % 32 != 0, block_log, empty problem
    assert lib.bq_rope_quantize(None, None, None, None, None, 64, 2, 64, 4, 64, 256, 256, ctypes.byref(f6), ctypes.byref(f6), None, None, None) == 1
    assert lib.bq_rope_quantize(256, 256, 256, 256, None, 64, 2, 40, 4, 64, 256, 256, ctypes.byref(f6), ctypes.byref(f6), 256, 256, None) == 2
    assert lib.bq_rope_quantize(256, 256, 256, 256, None, 64, 2, 64, 4, 48, 192, 192, ctypes.byref(f6), ctypes.byref(f6), 256, 256, None) == 2
    assert lib.bq_rope_quantize(256, 256, 256, 256, None, 64, 2, 64, 4, 64, 256, 256, ctypes.byref(fl), ctypes.byref(f6), 256, 256, None) == 2
    assert lib.bq_rope_quantize(256, 256, 256, 256, None, 32, 2, 64, 4, 64, 256, 256, ctypes.byref(f6), ctypes.byref(f6), 256, 256, None) == 1  # table shorter than S
    assert lib.bq_rope_quantize(None, None, None, None, None, 0, 0, 64, 4, 64, 256, 256, ctypes.byref(f6), ctypes.byref(f6), None, None, None) == 0
    # batched split GEMM: scales are mandatory, K % 8, batch stride smaller than one matrix, empty batch
    ta = (ctypes.c_int32 * 2)(0, 0)
    tb = (ctypes.c_int32 * 2)(1, 0)
    assert lib.bq_bmm_split16_tn(256, 256, 256, None, None, 2, 8, 8, 8, 2, ta, tb, 8, 64, None) == 1
    assert lib.bq_bmm_split16_tn(256, 256, 256, 256, 256, 2, 8, 8, 12, 2, ta, tb, 8, 64, None) == 1
    assert lib.bq_bmm_split16_tn(256, 256, 256, 256, 256, 2, 8, 8, 8, 2, ta, tb, 8, 32, None) == 1
    assert lib.bq_bmm_split16_tn(256, 256, 256, 256, 256, 0, 8, 8, 8, 2, ta, tb, 8, 64, None) == 0
    # norm + quantize: H % 16, H beyond one block per thread, unsupported kind, more than three outputs
    fm = (L.BqFormat * 1)(f6)
    outs = (ctypes.c_void_p * 1)(256)
    assert lib.bq_norm_quantize(256, 4, 40, 40, 256, None, 1e-5, 1, fm, outs, None) == 2
    assert lib.bq_norm_quantize(256, 4, 16 * 513, 16 * 513, 256, None, 1e-5, 1, fm, outs, None) == 2
    fd = L.BqFormat(L.KIND["minifloat_denorm"], 8, 4, 7, 0, 1, 16, 0)
    assert lib.bq_norm_quantize(256, 4, 64, 64, 256, None, 1e-5, 1, (L.BqFormat * 1)(fd), outs, None) == 2
    # split-attention entry points (block_log path): format, geometry and pointer checks precede any CUDA call
    assert lib.bq_softmax_quantize(ctypes.byref(fd), 256, 256, 4, 2, 64, 64, 64, 4096, 64, 4096, 1.0, 1, None, 0, None) == 2
    assert lib.bq_softmax_quantize(ctypes.byref(fl), 256, 256, 4, 2, 64, 40, 40, 2560, 40, 2560, 1.0, 0, None, 0, None) == 2     # Sk % 16
    assert lib.bq_softmax_quantize(ctypes.byref(fl), 256, 256, 4, 2, 64, 128, 128, 8192, 128, 8192, 1.0, 1, None, 0, None) == 2  # causal needs Sq == Sk
    assert lib.bq_softmax_quantize(ctypes.byref(fl), 256, 256, 3, 2, 64, 64, 64, 4096, 64, 4096, 1.0, 1, None, 0, None) == 1     # batch % heads
    assert lib.bq_softmax_quantize(ctypes.byref(fl), 256, 256, 4, 2, 64, 64, 64, 4096, 64, 4096, 0.0, 1, None, 0, None) == 1     # score_div
    assert lib.bq_softmax_quantize(ctypes.byref(fl), 256, 256, 4, 2, 64, 64, 64, 4096, 64, 4096, 1.0, 1, 256, 1, None) == 1      # bitmap too short
    assert lib.bq_softmax_quantize(ctypes.byref(fl), None, None, 0, 2, 64, 64, 64, 4096, 64, 4096, 1.0, 1, None, 0, None) == 0
    assert lib.bq_rope_quantize_split(256, 256, 256, None, None, 64, 2, 64, 4, 64, 256, 256, ctypes.byref(fl), 256, 256, None) == 1   # one table only
    assert lib.bq_rope_quantize_split(256, 256, 256, 256, None, 64, 2, 64, 4, 48, 192, 192, ctypes.byref(fl), 256, 256, None) == 2  # head_dim % 32
    assert lib.bq_rope_quantize_split(256, 256, 256, 256, None, 32, 2, 64, 4, 64, 256, 256, ctypes.byref(fl), 256, 256, None) == 1  # table shorter than S
    assert lib.bq_rope_quantize_split(256, 256, None, None, None, 0, 2, 64, 4, 64, 256, 256, ctypes.byref(fd), 256, 256, None) == 2 # format
    assert lib.bq_split3_bf16_transposed(256, 256, 2, 63, 4, 64, 256, None) == 2
    assert lib.bq_split3_bf16_transposed(256, 256, 2, 64, 4, 64, 128, None) == 1
    t3a = (ctypes.c_int32 * 3)(0, 0, 0)
    t3b = (ctypes.c_int32 * 3)(2, 1, 0)
    assert lib.bq_bmm_split_tn(256, 256, 256, 4, 64, 64, 64, 1, 3, 3, t3a, t3b, 64, 32, 0, None) == 1        # C batches would overlap
    assert lib.bq_bmm_split_tn(256, 256, 256, 4, 64, 64, 64, 1, 2, 3, t3a, t3b, 64, 4096, 0, None) == 1      # term names a missing plane
    assert lib.bq_bmm_split_tn(256, 256, 256, 4, 64, 128, 64, 1, 3, 3, t3a, t3b, 128, 8192, 1, None) == 1   # causal 1 needs M == N
    assert lib.bq_bmm_split_tn(256, 256, 256, 0, 64, 64, 64, 1, 3, 3, t3a, t3b, 64, 4096, 0, None) == 0
    assert lib.bq_norm_quantize(256, 4, 64, 64, 256, None, 1e-5, 4, fm, outs, None) == 1
    assert lib.bq_norm_quantize(256, 0, 64, 64, 256, None, 1e-5, 1, fm, outs, None) == 0
    # GEMM with the RoPE + quantizer epilogue (A, B, C, bias, fmt, qdir, cos, sin, pos, table_rows, S, head_dim, M, N, K, lda, ldb, ldc)
    rope = lambda *a: lib.bq_gemm_bf16_tn_rope(*a, None)
    assert rope(256, 256, 256, None, None, 0, 256, 256, None, 64, 64, 128, 64, 256, 64, 64, 64, 256) == 1                      # no format
    assert rope(256, 256, 256, None, ctypes.byref(f6), 0, None, 256, None, 64, 64, 128, 64, 256, 64, 64, 64, 256) == 1         # no cos table
    assert rope(256, 256, 256, None, ctypes.byref(f6), 0, 256, 256, None, 64, 64, 96, 64, 192, 64, 64, 64, 192) == 2           # head_dim 96
    assert rope(256, 256, 256, None, ctypes.byref(f6), 0, 256, 256, None, 64, 64, 128, 64, 192, 64, 64, 64, 192) == 2          # N % head_dim
    assert rope(256, 256, 256, None, ctypes.byref(fl), 0, 256, 256, None, 64, 64, 128, 64, 256, 64, 64, 64, 256) == 2          # block_log
    assert rope(256, 256, 256, None, ctypes.byref(f6), 1, 256, 256, None, 64, 64, 128, 72, 256, 64, 64, 64, 256) == 2          # qdir 1: M % 16
    assert rope(256, 256, 256, None, ctypes.byref(f6), 0, 256, 256, None, 32, 64, 128, 64, 256, 64, 64, 64, 256) == 1          # table shorter than S
    assert rope(256, 256, 256, None, ctypes.byref(f6), 0, 256, 256, None, 64, 64, 128, 64, 256, 64, 64, 64, 128) == 1          # ldc < N
    assert rope(None, None, None, None, ctypes.byref(f6), 0, None, None, None, 64, 64, 128, 0, 256, 64, 64, 64, 256) == 0      # empty problem
    # q | k | v in one launch: H % 256, M % 16, missing output / format
    qkv = lambda *a: lib.bq_gemm_bf16_tn_qkv_rope(*a, None)
    F6 = ctypes.byref(f6)
    assert qkv(256, 256, 256, 256, None, None, F6, F6, F6, 256, 256, None, 64, 64, 128, 64, 256, 64, 64, 64, 256) == 1          # no Cv
    assert qkv(256, 256, 256, 256, 256, None, F6, None, F6, 256, 256, None, 64, 64, 128, 64, 256, 64, 64, 64, 256) == 1         # no fk
    assert qkv(256, 256, 256, 256, 256, None, F6, F6, F6, 256, 256, None, 64, 64, 128, 64, 384, 64, 64, 64, 384) == 2           # H % 256
    assert qkv(256, 256, 256, 256, 256, None, F6, F6, F6, 256, 256, None, 64, 64, 128, 72, 256, 64, 64, 64, 256) == 2           # M % 16
    assert qkv(256, 256, 256, 256, 256, None, F6, F6, ctypes.byref(fl), 256, 256, None, 64, 64, 128, 64, 256, 64, 64, 64, 256) == 2   # block_log v
    assert qkv(None, None, None, None, None, None, F6, F6, F6, None, None, None, 64, 64, 128, 0, 256, 64, 64, 64, 256) == 0     # empty
    # gated-SiLU epilogue (act = 2): needs a format along N, bf16 output, no residual / replicas, ldc >= N / 2
    ep = L.BqGemmEpilogue()
    ep.scale, ep.act, ep.out_dtype = 1.0, 2, L.BQ_BF16
    ex = lambda ldc=128: lib.bq_gemm_bf16_tn_ex(256, 256, 256, ctypes.byref(ep), 32, 256, 64, 64, 64, ldc, None)
    assert ex() == 2                                                                              # no qfmt
    ep.qfmt = ctypes.pointer(f6)
    assert ex(64) == 1                                                                            # ldc < N / 2
    ep.qdir = 1
    assert ex() == 2                                                                              # blocks along M
    ep.qdir, ep.out_dtype = 0, L.BQ_F32
    assert ex() == 2                                                                              # fp32 output
    ep.out_dtype, ep.scale = L.BQ_BF16, 0.5
    assert ex() == 2                                                                              # scale
    ep.scale, ep.act = 1.0, 3
    assert ex(256) == 2                                                                           # unknown activation
    del q


def test_no_cpu_fallback():
    import torch

    from llm_mixed_q_b200.models.quantize.quantizers import block_fp_quantizer

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        block_fp_quantizer(torch.randn(4, 32), 6, 8, 127, [1, 16], True)
