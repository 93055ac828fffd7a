"""
Host side of block()/unblock() (reference quantizers/utils.py:42-321).

The reference materialises blocks with F.pad + F.unfold and undoes it with F.fold.  Here blocking
is pure index math inside the kernels; the host only has to (1) resolve the block shape the
reference would infer, (2) canonicalise the four layouts the reference distinguishes (1-D bias,
2-D activation, 2-D weight, 3-D activation) to one strided 3-D problem [L, R, C] with a
(b0, b1) block over (R, C), and (3) launch.  Exceptions mirror the reference's.
"""
from __future__ import annotations

import ctypes

import torch
from torch import Tensor

from .... import _lib as L


def resolve_block_shape(x_shape, block_size):
    """Same answers as `_infer_block_shape` (reference utils.py:42-67): right-align, -1/oversize -> whole dim."""
    dims = [int(d) for d in x_shape]
    blk = [int(b) for b in block_size]
    n = len(dims)
    blk = blk[len(blk) - n:] if len(blk) >= n else [-1] * (n - len(blk)) + blk
    return [d if (b == -1 or b > d) else b for d, b in zip(dims, blk)]


class Canon:
    """[L, R, C] view of a quantizer operand + block extents + whether the reference folds (−0.0 → +0.0)."""

    __slots__ = ("L", "R", "C", "sL", "sR", "sC", "b0", "b1", "fold", "shape")

    def desc(self) -> L.BqTensor3:
        return L.BqTensor3(self.L, self.R, self.C, self.sL, self.sR, self.sC)


def canonicalise(x: Tensor, block_size, skip_first_dim: bool, blocked: bool = True) -> Canon:
    c = Canon()
    c.shape = tuple(x.shape)
    if isinstance(block_size, int):
        block_size = [block_size]
    st = x.stride()
    if not blocked:
        # element-wise formats: any shape; flatten to one row when dense, else fall back to a 3-D view
        xc = x if x.is_contiguous() else None
        if xc is not None or x.ndim <= 1:
            n = x.numel()
            c.L, c.R, c.C, c.sL, c.sR, c.sC = 1, 1, n, n, n, (st[0] if x.ndim == 1 else 1)
            c.b0 = c.b1 = 1
            c.fold = False
            return c
        raise _NeedsContiguous()
    if x.ndim == 1:
        assert skip_first_dim is False, "skip_first_dim must be False for bias to be blocked"
        (b,) = resolve_block_shape(x.shape, block_size)
        c.L, c.R, c.C, c.sL, c.sR, c.sC = 1, 1, x.shape[0], 0, 0, st[0]
        c.b0, c.b1, c.fold = 1, b, False
    elif x.ndim == 2:
        if skip_first_dim:
            bs = resolve_block_shape([1, x.shape[1]], block_size)
            c.L, c.R, c.C, c.sL, c.sR, c.sC = x.shape[0], 1, x.shape[1], st[0], 0, st[1]
            c.b0, c.b1, c.fold = 1, bs[1], False
        else:
            bs = resolve_block_shape(x.shape, block_size)
            c.L, c.R, c.C, c.sL, c.sR, c.sC = 1, x.shape[0], x.shape[1], 0, st[0], st[1]
            c.b0, c.b1, c.fold = bs[0], bs[1], True
    elif x.ndim == 3:
        if not skip_first_dim:
            raise NotImplementedError("block 3d weight is not supported.")
        bs = resolve_block_shape([1, x.shape[1], x.shape[2]], block_size)
        c.L, c.R, c.C = x.shape
        c.sL, c.sR, c.sC = st
        c.b0, c.b1, c.fold = bs[1], bs[2], True
    else:
        raise RuntimeError(f"Unsupported x.ndim = {x.ndim}")
    c.b0, c.b1 = max(int(c.b0), 1), max(int(c.b1), 1)
    return c


class _NeedsContiguous(Exception):
    pass


def make_format(kind: str, *, width=0, exponent_width=0, exponent_bias=0, exponent_bias_width=0, b0=1, b1=1, fold=False):
    return L.BqFormat(L.KIND[kind], int(width), int(exponent_width), int(exponent_bias), int(exponent_bias_width), int(b0),
                      int(b1), 1 if fold else 0)


def default_bias(exponent_bias, exponent_width):
    """reference block_fp.py:61-62 / minifloat.py:51-52: None / "none" / "None" -> 2^(ew-1) - 1."""
    if exponent_bias in (None, "none", "None"):
        return 2 ** (int(exponent_width) - 1) - 1
    return exponent_bias


def launch_quantize(x: Tensor, fmt: L.BqFormat, canon: Canon, out_dtype=torch.float32, transpose_out=False, out=None):
    """One bq_quantize call.  Returns a new tensor of the operand's logical shape (or [L, C, R] if transposed)."""
    lib = L.load()
    if canon.L * canon.R * canon.C == 0:
        return torch.empty(canon.shape, dtype=out_dtype, device=x.device)
    desc = canon.desc()
    if out is None:
        oshape = (canon.L, canon.C, canon.R) if transpose_out else canon.shape
        out = torch.empty(oshape, dtype=out_dtype, device=x.device)
    nbytes = lib.bq_quantize_workspace_bytes(ctypes.byref(fmt), ctypes.byref(desc))
    ws = L.workspace(nbytes, x.device)
    rc = lib.bq_quantize(ctypes.byref(fmt), ctypes.byref(desc), x.data_ptr(), out.data_ptr(),
                         L.BQ_F32 if out_dtype == torch.float32 else L.BQ_BF16, 1 if transpose_out else 0, ws.data_ptr(),
                         ws.numel(), L.stream_ptr(x.device))
    L.check(rc, "bq_quantize")
    return out


class _STE(torch.autograd.Function):
    """Straight-through backward, as every reference quantizer (e.g. block_fp.py:119-124)."""

    @staticmethod
    def forward(ctx, x, fn):
        return fn(x)

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output, None


def with_ste(x: Tensor, fn):
    if torch.is_grad_enabled() and x.requires_grad:
        return _STE.apply(x, fn)
    return fn(x)


def quantize_blocked(x: Tensor, kind: str, block_size, skip_first_dim: bool, **fmt_kw):
    L.require_cuda_f32(x, "x")
    canon = canonicalise(x.detach(), block_size, skip_first_dim, blocked=True)
    fmt = make_format(kind, b0=canon.b0, b1=canon.b1, fold=canon.fold, **fmt_kw)
    return with_ste(x, lambda t: launch_quantize(t.detach(), fmt, canon))


def quantize_elementwise(x: Tensor, kind: str, **fmt_kw):
    L.require_cuda_f32(x, "x")
    xd = x.detach()
    try:
        canon = canonicalise(xd, None, False, blocked=False)
    except _NeedsContiguous:
        xd = xd.contiguous()
        canon = canonicalise(xd, None, False, blocked=False)
    fmt = make_format(kind, **fmt_kw)
    src = xd
    return with_ste(x, lambda t: launch_quantize(src, fmt, canon))
