"""
Glue fusion between the quantized ops of a decoder layer — host side of bq_norm_quantize and of
the quantising GEMM / attention epilogues (include/bq.h).

The reference runs every op of a layer as its own torch kernels and round-trips fp32 activations through HBM between
them (models/opt_quantized/modeling_opt.py:360-441).  In PTQ inference the x-quantizer of an op only depends on the
tensor it reads, so it can run inside the kernel that PRODUCES that tensor and hand the next GEMM a bf16 operand
holding the exact quantised values.  These helpers decide when that is legal and build the format descriptors; the
arithmetic (and its op order) is the reference's.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence, Tuple

import torch

from .... import _lib as L
from ..quantized_modules.linear import _LinearBase, operand_format, significant_bits
from ..quantizers.utils import canonicalise, make_format, resolve_block_shape

_EPI_KINDS = ("block_fp", "block_minifloat")


_NORM_KINDS = ("block_fp", "block_minifloat", "block_log")    # block_log: carrier rule of include/bq.h (block-local, bf16)


def row_block16_format(config: dict, prefix: str, last_dim: int, rows: int = 1, kinds=_EPI_KINDS) -> Optional[Tuple[str, dict]]:
    """(kind, kwargs) when `<prefix>_*` of a config node is a block_fp / block_minifloat format that resolves to blocks of
    16 along a last dim of size `last_dim` and is exact in bf16 — i.e. something an epilogue can apply — else None.
    `rows`: second-to-last extent of the operand THE REFERENCE blocks (quantizers/utils.py:42-67 right-aligns the block
    size and fills missing leading entries with -1 = whole dim): S for a Linear fed a 3-D [B, S, H] activation or a
    [B*h, S, d] bmm operand, 1 for a Linear fed a 2-D activation (skip_first_dim drops the row dim, utils.py:127-144).  With
    `block_size = [16]` a 3-D operand resolves to [1, S, 16] — a block spans every token — which no epilogue implements."""
    try:
        if config is None or config.get("bypass", False):
            return None
        kind, kw, bs = operand_format(config, prefix)
    except KeyError:
        return None
    if kind not in kinds or bs is None or significant_bits(kind, kw) > 8:
        return None
    b = resolve_block_shape([1, max(int(rows), 1), last_dim], bs)
    if b[1] != 1 or b[2] != 16 or last_dim % 16:
        return None
    return kind, kw


def linear_input_format(lin, last_dim: Optional[int] = None, rows: int = 1, kinds=_EPI_KINDS) -> Optional[Tuple[str, dict]]:
    """x-quantizer of a quantized Linear as an epilogue format, if the module can take a pre-quantised bf16 input.
    `rows` = S when the reference feeds the Linear a 3-D [B, S, K] activation, 1 when it feeds a 2-D one."""
    if not isinstance(lin, _LinearBase) or not lin.accepts_prequantized():
        return None
    return row_block16_format(lin.config, "data_in", lin.in_features if last_dim is None else last_dim, rows, kinds)


def _same(a, b) -> bool:
    return a[0] == b[0] and a[1] == b[1]


@torch.no_grad()
def norm_quantize(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], eps: float,
                  formats: Sequence[Tuple[str, dict]]) -> List[torch.Tensor]:
    """
    LayerNorm (bias given) or RMSNorm (bias None) of fp32 x [..., H], then one bf16 tensor per entry of `formats` holding
    Q_format(norm(x)); identical formats share one output tensor (q/k/v_proj usually do).
    """
    lib = L.load()
    H = x.shape[-1]
    x2 = x.reshape(-1, H)
    if x2.stride(-1) != 1 or (x2.shape[0] > 1 and x2.stride(0) % 4 != 0):
        x2 = x2.contiguous()
    rows = x2.shape[0]
    distinct: List[Tuple[str, dict]] = []
    index = []
    for f in formats:
        for i, d in enumerate(distinct):
            if _same(f, d):
                index.append(i)
                break
        else:
            index.append(len(distinct))
            distinct.append(f)
    outs = [torch.empty((rows, H), dtype=torch.bfloat16, device=x.device) for _ in distinct]
    for lo in range(0, len(distinct), 3):
        chunk = distinct[lo:lo + 3]
        fmts = (L.BqFormat * len(chunk))(*[make_format(k, b0=1, b1=16, **kw) for k, kw in chunk])
        ptrs = (ctypes.c_void_p * len(chunk))(*[o.data_ptr() for o in outs[lo:lo + 3]])
        rc = lib.bq_norm_quantize(x2.data_ptr(), rows, H, x2.stride(0) if rows > 1 else H, weight.data_ptr(),
                                  bias.data_ptr() if bias is not None else None, float(eps), len(chunk), fmts, ptrs,
                                  L.stream_ptr(x.device))
        L.check(rc, "bq_norm_quantize")
    shape = tuple(x.shape)
    return [outs[i].view(shape) for i in index]


@torch.no_grad()
def silu_mul_quantize(gate: torch.Tensor, up: torch.Tensor, fmt: Tuple[str, dict], out_dtype=torch.bfloat16) -> torch.Tensor:
    """Q_fmt(silu(gate) * up) for fp32 [rows, I] gate / up of identical layout, blocks [1,16] along I — the operand of Llama's
    down_proj (reference modeling_llama.py:246 + the x-quantizer of quantized_modules/linear.py:63-71) in one kernel:
    10 B/element of HBM traffic instead of 26 for silu, mul and quantize run separately."""
    lib = L.load()
    kind, kw = fmt
    assert gate.shape == up.shape and gate.ndim == 2 and gate.dtype == torch.float32 and up.dtype == torch.float32
    if not gate.is_contiguous():
        gate = gate.contiguous()
    if not up.is_contiguous():
        up = up.contiguous()
    canon = canonicalise(gate, [1, 16], True, blocked=True)
    f = make_format(kind, b0=1, b1=canon.b1, **kw)
    desc = canon.desc()
    out = torch.empty(gate.shape, dtype=out_dtype, device=gate.device)
    if gate.numel() == 0:
        return out
    ws = L.workspace(lib.bq_quantize_workspace_bytes(ctypes.byref(f), ctypes.byref(desc)), gate.device)
    rc = lib.bq_silu_mul_quantize(ctypes.byref(f), ctypes.byref(desc), gate.data_ptr(), up.data_ptr(), out.data_ptr(),
                                  L.BQ_F32 if out_dtype == torch.float32 else L.BQ_BF16, ws.data_ptr(), ws.numel(),
                                  L.stream_ptr(gate.device))
    L.check(rc, "bq_silu_mul_quantize")
    return out
