"""
fp32-equivalent Linear on the 16-bit tensor cores (host side of bq_split2_f16_rows / bq_gemm_split16_tn and of the older
bq_split3_bf16 / bq_gemm_split_tn).

The reference keeps a few matmuls UNQUANTISED in fp32 — above all the lm_head
(models/opt_quantized/modeling_opt.py:942-944, models/llama_quantized/modeling_llama.py:772), a plain
`nn.Linear` that costs 3.4 of the 49.6 TFLOP of an OPT-1.3B forward and would otherwise run on the fp32 SIMT pipe.
Each fp32 operand row is scaled by a power of two and split into two fp16 planes (x * 2^e = hi + lo, 22 significant
bits); the products lo*hi, hi*lo, hi*hi are accumulated in fp32 on tcgen05 and the scales are undone exactly in the
epilogue.  Per-product relative error ~2^-21: below the accumulation-order noise sqrt(K) * 2^-24 * sum|a||b| of an fp32
GEMM (tests/test_gpu_consumers.py asserts the same bound for both).  `mode="bf16x3"` selects the 6-term bf16 split
(2^-24 per product, twice the tensor work).
"""
from __future__ import annotations

import ctypes

import torch
import torch.nn.functional as F

from .... import _lib as L

# (plane of x, plane of w), smallest magnitude first so small terms are not absorbed by the accumulator
_TERMS6 = [(2, 0), (0, 2), (1, 1), (1, 0), (0, 1), (0, 0)]
_TERMS3 = [(1, 0), (0, 1), (0, 0)]
_weight_planes: dict = {}
MODE = "f16x2"


def split3(x: torch.Tensor) -> torch.Tensor:
    """fp32 [..] (contiguous, numel % 4 == 0) -> bf16 [3, ..]."""
    lib = L.load()
    out = torch.empty((3,) + tuple(x.shape), dtype=torch.bfloat16, device=x.device)
    L.check(lib.bq_split3_bf16(x.data_ptr(), out.data_ptr(), x.numel(), L.stream_ptr(x.device)), "bq_split3_bf16")
    return out


def split2_rows(x2: torch.Tensor):
    """fp32 [rows, K] (unit stride along K) -> (fp16 planes [2, rows, K], inv_scale fp32 [rows])."""
    lib = L.load()
    rows, K = x2.shape
    planes = torch.empty((2, rows, K), dtype=torch.float16, device=x2.device)
    inv = torch.empty((rows,), dtype=torch.float32, device=x2.device)
    L.check(lib.bq_split2_f16_rows(x2.data_ptr(), rows, K, x2.stride(0) if rows > 1 else K, planes.data_ptr(), inv.data_ptr(),
                                   L.stream_ptr(x2.device)), "bq_split2_f16_rows")
    return planes, inv


def _cached_weight(w: torch.Tensor, mode: str):
    key = (w.data_ptr(), w._version, w.device, tuple(w.shape), mode)
    hit = _weight_planes.get(id(w))
    if hit is None or hit[0] != key:
        wc = w.detach().contiguous()
        hit = (key, split2_rows(wc) if mode == "f16x2" else split3(wc))
        _weight_planes[id(w)] = hit
    return hit[1]


def fp32_linear(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None = None, mode: str | None = None) -> torch.Tensor:
    """F.linear(x, weight, bias) with fp32-equivalent accuracy; falls back to F.linear off the fast path."""
    mode = MODE if mode is None else mode
    K = x.shape[-1]
    if (not x.is_cuda or x.dtype != torch.float32 or weight.dtype != torch.float32 or K % 8 != 0
            or (torch.is_grad_enabled() and (x.requires_grad or weight.requires_grad))):
        return F.linear(x, weight, bias)
    lib = L.load()
    x2 = x.reshape(-1, K)
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    M, N = x2.shape[0], weight.shape[0]
    if M == 0:
        return F.linear(x, weight, bias)
    y = torch.empty((M, N), dtype=torch.float32, device=x.device)
    b = bias.detach() if bias is not None else None
    if mode == "f16x2":
        xp, xs = split2_rows(x2)
        wp, ws = _cached_weight(weight, mode)
        ta = (ctypes.c_int32 * 3)(*[t[0] for t in _TERMS3])
        tb = (ctypes.c_int32 * 3)(*[t[1] for t in _TERMS3])
        rc = lib.bq_gemm_split16_tn(xp.data_ptr(), wp.data_ptr(), y.data_ptr(), b.data_ptr() if b is not None else None,
                                    xs.data_ptr(), ws.data_ptr(), M, N, K, 3, ta, tb, N, L.stream_ptr(x.device))
        L.check(rc, "bq_gemm_split16_tn")
    else:
        xp = split3(x2)
        wp = _cached_weight(weight, mode)
        ta = (ctypes.c_int32 * 6)(*[t[0] for t in _TERMS6])
        tb = (ctypes.c_int32 * 6)(*[t[1] for t in _TERMS6])
        rc = lib.bq_gemm_split_tn(xp.data_ptr(), wp.data_ptr(), y.data_ptr(), b.data_ptr() if b is not None else None, M, N, K, 3, 3,
                                  6, ta, tb, N, L.stream_ptr(x.device))
        L.check(rc, "bq_gemm_split_tn")
    return y.reshape(*x.shape[:-1], N)
