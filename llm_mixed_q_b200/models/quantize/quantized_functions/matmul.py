"""
Quantized matmul / bmm — host mirror of reference quantized_functions/matmul.py:25-353.

`fn(x, y, config)` with fn from QUANTIZED_FUNC_MAP["matmul" | "bmm"][config["name"]]; `data_in_*` keys describe
x, `weight_*` keys describe y.  Reference semantics that are kept:
  * operands with more than 2 dims are flattened to 3-D over the leading dims and blocked with
    skip_first_dim=True, 2-D operands with skip_first_dim=False (matmul.py:165-193);
  * x is blocked along its last dim (the contraction dim), y along ITS last dim (the output-N dim) — for
    bmm_0 y is the k^T view, so its blocks run over 16 consecutive key positions (modeling_opt.py:246);
  * block_log (and its alias "log") quantises x only, y stays fp32 (matmul.py:293-296).

Underneath: one bq_bmm call = quantize x -> bf16, quantize y -> bf16 written K-major, tcgen05 GEMM.
"""
from __future__ import annotations

import ctypes

import torch

from .... import _lib as L
from ..quantized_modules.linear import operand_format, quantize_operand_bf16, significant_bits
from ..quantizers import QUANTIZER_MAP
from ..quantizers.utils import canonicalise, make_format

matmul_mapping = {"matmul": torch.matmul, "bmm": torch.bmm}


def _quantize_fp32(t, kind, kw, block_size, skip_first_dim):
    fn = QUANTIZER_MAP[kind]
    if kind in ("minifloat_denorm", "minifloat_ieee"):
        return fn(t, width=kw["width"], exponent_width=kw["exponent_width"], exponent_bias=kw["exponent_bias"])
    if kind == "integer":
        return fn(t, width=kw["width"], frac_width=kw["exponent_bias"])
    return fn(t, block_size=block_size, skip_first_dim=skip_first_dim, **kw)


def _flatten3(t):
    return torch.flatten(t, 0, -3) if t.ndim > 2 else t


def _generic_matmul(x, y, config, style, quantize_y=True):
    matmul = matmul_mapping[style]
    if config.get("bypass", False):
        return matmul(x, y)
    xk, xkw, xbs = operand_format(config, "data_in")
    yk, ykw, ybs = operand_format(config, "weight")
    if style == "bmm" and (x.ndim != 3 or y.ndim != 3):
        raise RuntimeError("batch1 must be a 3D tensor" if x.ndim != 3 else "batch2 must be a 3D tensor")
    x_multi, y_multi = x.ndim > 2, y.ndim > 2

    fusable = (
        x.is_cuda and y.is_cuda and x.dtype == torch.float32 and y.dtype == torch.float32
        and quantize_y and x.ndim >= 2 and y.ndim >= 2
        and significant_bits(xk, xkw) <= 8 and significant_bits(yk, ykw) <= 8
        and x.shape[-1] % 8 == 0 and x.shape[-1] == y.shape[-2]
        and (not torch.is_grad_enabled() or not (x.requires_grad or y.requires_grad))
        and (x.shape[:-2] == y.shape[:-2] or y.ndim == 2)
    )
    if fusable:
        x3, y3 = _flatten3(x), _flatten3(y)
        if x3.ndim == 2:
            x3 = x3.unsqueeze(0)
        xb0 = yb0 = 1
        if xbs is not None:
            xb0 = canonicalise(x3, xbs, True).b0 if x_multi else canonicalise(x3[0], xbs, False).b0
            yc = canonicalise(y3, ybs, True) if y_multi else canonicalise(y3, ybs, False)
            yb0 = yc.b0
        if xb0 == 1 and yb0 == 1:
            out = _fused_bmm(x3, y3 if y3.ndim == 3 else y3.unsqueeze(0), config, xk, xkw, xbs, yk, ykw, ybs)
            return out.reshape(*x.shape[:-1], y.shape[-1])

    # general route: reference op order with our quantizer kernels, then an fp32(-equivalent) matmul
    x_shape, y_shape = list(x.shape), list(y.shape)
    xq = _quantize_fp32(_flatten3(x), xk, xkw, xbs, x_multi).reshape(x_shape)
    yq = _quantize_fp32(_flatten3(y), yk, ykw, ybs, y_multi).reshape(y_shape) if quantize_y else y
    if (SPLIT_BMM and xq.is_cuda and yq.is_cuda and xq.dtype == torch.float32 and yq.dtype == torch.float32
            and xq.ndim >= 2 and yq.ndim >= 2 and xq.shape[-1] == yq.shape[-2] and xq.shape[-1] % 8 == 0
            and (xq.shape[:-2] == yq.shape[:-2] or yq.ndim == 2) and xq.numel() > 0 and yq.numel() > 0
            and (not torch.is_grad_enabled() or not (xq.requires_grad or yq.requires_grad))):
        # x is a power of two under block_log: its low fp16 plane is identically zero, two products suffice
        return _split_bmm(xq, yq, x_lo_is_zero=(xk == "block_log"))
    return matmul(xq, yq)


SPLIT_BMM = True        # False: fp32 library matmul on the general route (A/B measurement, tests)


def _split_bmm(xq, yq, x_lo_is_zero=False):
    """matmul(xq, yq) for fp32 operands that are not bf16-exact, on the tensor cores: every row of xq and every column of yq is
    scaled by a power of two and split into two fp16 planes (bq_split2_f16_rows), the plane products are accumulated in fp32
    (bq_bmm_split16_tn) — ~2^-21 relative error per product, inside the accumulation-order noise of the fp32 matmul the reference
    runs here (quantized_functions/matmul.py:196, :293-296).  Stands in for the SIMT SGEMM torch.matmul would launch."""
    from .fp32_linear import split2_rows

    lib = L.load()
    out_shape = tuple(xq.shape[:-1]) + (yq.shape[-1],)
    x3 = _flatten3(xq)
    if x3.ndim == 2:
        x3 = x3.unsqueeze(0)
    batch, M, K = x3.shape
    N = yq.shape[-1]
    y3 = _flatten3(yq)
    if y3.ndim == 2:                                    # one weight-like matrix for every batch: fold the batch into M
        x3 = x3.reshape(1, batch * M, K)
        y3 = y3.unsqueeze(0)
        batch, M = 1, batch * M
    yt = y3.transpose(1, 2).contiguous()               # [batch, N, K]; a no-op for the k^T views the attention passes
    ap, a_inv = split2_rows(x3.reshape(batch * M, K).contiguous())
    bp, b_inv = split2_rows(yt.reshape(batch * N, K))
    out = torch.empty((batch, M, N), dtype=torch.float32, device=xq.device)
    terms = [(0, 1), (0, 0)] if x_lo_is_zero else [(1, 0), (0, 1), (0, 0)]        # smallest magnitude first
    ta = (ctypes.c_int32 * len(terms))(*[t[0] for t in terms])
    tb = (ctypes.c_int32 * len(terms))(*[t[1] for t in terms])
    rc = lib.bq_bmm_split16_tn(ap.data_ptr(), bp.data_ptr(), out.data_ptr(), a_inv.data_ptr(), b_inv.data_ptr(), batch, M, N, K,
                               len(terms), ta, tb, N, M * N, L.stream_ptr(xq.device))
    L.check(rc, "bq_bmm_split16_tn")
    return out.reshape(out_shape)


def _fused_bmm(x3, y3, config, xk, xkw, xbs, yk, ykw, ybs):
    lib = L.load()
    batch, M, K = x3.shape
    N = y3.shape[-1]
    if not x3.is_contiguous():
        x3 = x3.contiguous()
    if y3.shape[0] != batch:
        y3 = y3.expand(batch, K, N)
    sy = y3.stride()
    if not (sy[1] == 1 or sy[2] == 1):
        y3 = y3.contiguous()
        sy = y3.stride()
    xcols = canonicalise(x3, xbs, True).b1 if xbs is not None else 1
    ycols = canonicalise(y3, ybs, True).b1 if ybs is not None else 1
    fx = make_format(xk, b0=1, b1=xcols, **xkw)
    fy = make_format(yk, b0=1, b1=ycols, **ykw)
    out = torch.empty((batch, M, N), dtype=torch.float32, device=x3.device)
    if batch * M * N == 0:
        return out
    nbytes = lib.bq_bmm_workspace_bytes(ctypes.byref(fx), ctypes.byref(fy), batch, M, K, N)
    ws = L.workspace(nbytes, x3.device)
    rc = lib.bq_bmm(ctypes.byref(fx), ctypes.byref(fy), x3.data_ptr(), y3.data_ptr(), batch, M, K, N, sy[0], sy[1], sy[2],
                    out.data_ptr(), ws.data_ptr(), ws.numel(), L.stream_ptr(x3.device))
    L.check(rc, "bq_bmm")
    return out


def generic_matmul_block_fp(x, y, config, style="matmul"):
    return _generic_matmul(x, y, config, style)


def generic_matmul_block_minifloat(x, y, config, style="matmul"):
    return _generic_matmul(x, y, config, style)


def generic_matmul_block_log(x, y, config, style="matmul"):
    return _generic_matmul(x, y, config, style, quantize_y=False)


def generic_matmul_minifloat_denorm(x, y, config, style="matmul"):
    return _generic_matmul(x, y, config, style)


def generic_matmul_minifloat_ieee(x, y, config, style="matmul"):
    return _generic_matmul(x, y, config, style)


def generic_matmul_integer(x, y, config, style="matmul"):
    return _generic_matmul(x, y, config, style)


def matmul_block_fp(x, y, config):
    return generic_matmul_block_fp(x, y, config, "matmul")


def matmul_block_minifloat(x, y, config):
    return generic_matmul_block_minifloat(x, y, config, "matmul")


def matmul_block_log(x, y, config):
    return generic_matmul_block_log(x, y, config, "matmul")


def matmul_minifloat_denorm(x, y, config):
    return generic_matmul_minifloat_denorm(x, y, config, "matmul")


def matmul_minifloat_ieee(x, y, config):
    return generic_matmul_minifloat_ieee(x, y, config, "matmul")


def matmul_integer(x, y, config):
    return generic_matmul_integer(x, y, config, "matmul")


def bmm_block_fp(x, y, config):
    return generic_matmul_block_fp(x, y, config, "bmm")


def bmm_block_minifloat(x, y, config):
    return generic_matmul_block_minifloat(x, y, config, "bmm")


def bmm_block_log(x, y, config):
    return generic_matmul_block_log(x, y, config, "bmm")


def bmm_minifloat_denorm(x, y, config):
    return generic_matmul_minifloat_denorm(x, y, config, "bmm")


def bmm_minifloat_ieee(x, y, config):
    return generic_matmul_minifloat_ieee(x, y, config, "bmm")


def bmm_integer(x, y, config):
    return generic_matmul_integer(x, y, config, "bmm")
