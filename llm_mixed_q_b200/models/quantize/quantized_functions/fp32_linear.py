"""
fp32-equivalent Linear on the bf16 tensor cores (host side of bq_split3_bf16 / bq_gemm_split_tn).

The reference keeps a few matmuls UNQUANTISED in fp32 — above all the lm_head
(models/opt_quantized/modeling_opt.py:942-944, models/llama_quantized/modeling_llama.py:772), a plain
`nn.Linear` that costs 3.4 of the 49.6 TFLOP of an OPT-1.3B forward and would otherwise run on the fp32 SIMT
pipe.  Each fp32 operand is split error-free into three bf16 planes (x = x0 + x1 + x2); the six products whose
magnitude is >= 2^-24 of the leading one are accumulated in fp32 on tcgen05.  Per-product relative error is
~2^-24, the same order as fp32 rounding itself.
"""
from __future__ import annotations

import ctypes

import torch
import torch.nn.functional as F

from .... import _lib as L

# (plane of x, plane of w), smallest magnitude first so small terms are not absorbed by the accumulator
_TERMS6 = [(2, 0), (0, 2), (1, 1), (1, 0), (0, 1), (0, 0)]
_weight_planes: dict = {}


def split3(x: torch.Tensor) -> torch.Tensor:
    """fp32 [..] (contiguous, numel % 4 == 0) -> bf16 [3, ..]."""
    lib = L.load()
    out = torch.empty((3,) + tuple(x.shape), dtype=torch.bfloat16, device=x.device)
    L.check(lib.bq_split3_bf16(x.data_ptr(), out.data_ptr(), x.numel(), L.stream_ptr(x.device)), "bq_split3_bf16")
    return out


def _planes_of_weight(w: torch.Tensor) -> torch.Tensor:
    key = (w.data_ptr(), w._version, w.device, tuple(w.shape))
    hit = _weight_planes.get(id(w))
    if hit is None or hit[0] != key:
        hit = (key, split3(w.detach().contiguous()))
        _weight_planes[id(w)] = hit
    return hit[1]


def fp32_linear(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None = None) -> torch.Tensor:
    """F.linear(x, weight, bias) with fp32-equivalent accuracy; falls back to F.linear off the fast path."""
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
    xp = split3(x2)
    wp = _planes_of_weight(weight)
    y = torch.empty((M, N), dtype=torch.float32, device=x.device)
    ta = (ctypes.c_int32 * 6)(*[t[0] for t in _TERMS6])
    tb = (ctypes.c_int32 * 6)(*[t[1] for t in _TERMS6])
    b = bias.detach() if bias is not None else None
    rc = lib.bq_gemm_split_tn(xp.data_ptr(), wp.data_ptr(), y.data_ptr(), b.data_ptr() if b is not None else None, M, N, K, 3, 3,
                              6, ta, tb, N, L.stream_ptr(x.device))
    L.check(rc, "bq_gemm_split_tn")
    return y.reshape(*x.shape[:-1], N)
