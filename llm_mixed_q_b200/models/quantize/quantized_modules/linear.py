"""
Quantized Linear modules — host mirror of reference quantized_modules/linear.py:31-307.

Same class names, constructor (`cls(in_features, out_features, bias=True, device=None, dtype=None, config=cfg)`),
attributes (`config, bypass, is_ptq, weight_requires_quantisation, x_quantizer, w_quantizer, b_quantizer`),
`forward` semantics and `from_float` as the reference, so HF model code written against the reference's
`get_quantized_cls("linear", cfg)` runs unchanged.

What changes underneath (PTQ steady state, the path every shipped TOML selects):
  reference : x -> ~45 ATen kernels -> fp32 x_q -> cuBLAS SGEMM with the (in-place quantised) fp32 weight
  here      : x -> one streaming quantize kernel emitting bf16 -> tcgen05 GEMM against a bf16 cache of the
              quantised weight (exact: <= 8 significant bits), fp32 accumulation, fp32 bias add in the epilogue.
The in-place PTQ overwrite of `weight` / `bias` on the first forward (linear.py:66-70) is kept — the fp32
parameters hold bit-identical quantised values afterwards; the bf16 cache is derived state.
"""
from __future__ import annotations

import ctypes
from functools import partial

import torch
import torch.nn as nn
import torch.nn.functional as F

from .... import _lib as L
from ..quantizers import (block_fp_quantizer, block_log_quantizer, block_minifloat_quantizer, integer_quantizer,
                          minifloat_denorm_quantizer, minifloat_ieee_quantizer)
from ..quantizers.utils import canonicalise, default_bias, launch_quantize, make_format


def operand_format(config: dict, prefix: str):
    """(kind, fmt kwargs, block_size) of one operand from the `<prefix>_*` keys of a config node."""
    name = config["name"]
    if name == "block_fp":
        return "block_fp", dict(width=config[f"{prefix}_width"], exponent_width=config[f"{prefix}_exponent_width"],
                                exponent_bias=default_bias(config[f"{prefix}_exponent_bias"],
                                                           config[f"{prefix}_exponent_width"])), config[f"{prefix}_block_size"]
    if name == "block_minifloat":
        return "block_minifloat", dict(width=config[f"{prefix}_width"], exponent_width=config[f"{prefix}_exponent_width"],
                                       exponent_bias_width=config[f"{prefix}_exponent_bias_width"]), config[f"{prefix}_block_size"]
    if name in ("block_log", "log"):
        return "block_log", dict(width=config[f"{prefix}_width"],
                                 exponent_bias_width=config[f"{prefix}_exponent_bias_width"]), config[f"{prefix}_block_size"]
    if name in ("minifloat_denorm", "minifloat_ieee"):
        return name, dict(width=config[f"{prefix}_width"], exponent_width=config[f"{prefix}_exponent_width"],
                          exponent_bias=default_bias(config[f"{prefix}_exponent_bias"],
                                                     config[f"{prefix}_exponent_width"])), None
    if name == "integer":
        return "integer", dict(width=config[f"{prefix}_width"], exponent_bias=config[f"{prefix}_frac_width"]), None
    raise KeyError(name)


def significant_bits(kind: str, kw: dict) -> int:
    """Significant bits of a quantised value; <= 8 means the value is exact in bf16."""
    if kind == "block_fp":
        return kw["width"] - 1
    if kind in ("block_minifloat", "minifloat_ieee"):
        return kw["width"] - kw["exponent_width"]          # implicit one + mantissa
    if kind == "minifloat_denorm":
        return kw["width"] - kw["exponent_width"] - 1
    if kind == "block_log":
        return 1
    if kind == "integer":
        return kw["width"] - 1
    return 24


def quantize_operand_bf16(x: torch.Tensor, kind, kw, block_size, skip_first_dim, transpose_out=False):
    """Quantise to a bf16 GEMM operand (K-major); used for weights (once) and odd-shaped activations."""
    blocked = block_size is not None
    xd = x.detach()
    if not blocked and not xd.is_contiguous():
        xd = xd.contiguous()
    canon = canonicalise(xd, block_size, skip_first_dim, blocked=blocked)
    fmt = make_format(kind, b0=canon.b0, b1=canon.b1, fold=False, **kw)
    return launch_quantize(xd, fmt, canon, out_dtype=torch.bfloat16, transpose_out=transpose_out)


# A/B switches for the two north-star variants of the operator-API call (measured against the default two-launch path in
# DESIGN.md §3; both give bit-identical results to it):
#   FUSED_PROLOGUE   x-quantizer inside the GEMM prologue (bq_linear_fused): one launch, no bf16 copy of x in HBM
#   PACKED_WEIGHTS   weights held as w + 0.5 bits / element (sign+magnitude fields + one exponent byte per block of 16) and decoded to
#                    bf16 in the GEMM mainloop (bq_pack_weight / bq_gemm_packed_tn) instead of the 16-bit bf16 cache
FUSED_PROLOGUE = False
PACKED_WEIGHTS = False


def pack_weight(wq: torch.Tensor, width: int, exponent_width: int, exponent_bias: int):
    """Quantised fp32 weight [N, K] (on the block_fp grid, K % 256 == 0) -> (packed uint8 [N, K/256 * (32*width + 16)], number of
    elements the packed form does not reproduce bit for bit — the reference's pass-through elements)."""
    lib = L.load()
    L.require_cuda_f32(wq, "weight")
    N, K = wq.shape
    fmt = make_format("block_fp", width=width, exponent_width=exponent_width, exponent_bias=exponent_bias, b0=1, b1=16)
    nbytes = lib.bq_packed_weight_bytes(ctypes.byref(fmt), N, K)
    if nbytes == 0:
        raise NotImplementedError(f"packed weights need block_fp with 2 <= width <= 8 and K % 256 == 0 (width={width}, K={K})")
    w2 = wq.detach()
    if w2.stride(-1) != 1 or w2.stride(0) % 4:
        w2 = w2.contiguous()
    packed = torch.empty(nbytes, dtype=torch.uint8, device=wq.device)
    bad = torch.zeros(1, dtype=torch.int64, device=wq.device)
    L.check(lib.bq_pack_weight(ctypes.byref(fmt), w2.data_ptr(), N, K, w2.stride(0), packed.data_ptr(), bad.data_ptr(),
                               L.stream_ptr(wq.device)), "bq_pack_weight")
    return packed.view(N, -1), int(bad.item())


def gemm_packed(xq: torch.Tensor, packed: torch.Tensor, width: int, exponent_width: int, exponent_bias: int, N: int,
                bias: torch.Tensor = None) -> torch.Tensor:
    """y = xq @ unpack(packed)^T (+ bias): bf16 [M, K] activation operand against packed block_fp weights, decoded in the mainloop."""
    lib = L.load()
    K = xq.shape[-1]
    x2 = xq.reshape(-1, K)
    if x2.stride(-1) != 1 or (x2.shape[0] > 1 and x2.stride(0) % 8):
        x2 = x2.contiguous()
    M = x2.shape[0]
    fmt = make_format("block_fp", width=width, exponent_width=exponent_width, exponent_bias=exponent_bias, b0=1, b1=16)
    y = torch.empty((M, N), dtype=torch.float32, device=xq.device)
    if M:
        L.check(lib.bq_gemm_packed_tn(x2.data_ptr(), packed.data_ptr(), ctypes.byref(fmt), y.data_ptr(),
                                      bias.data_ptr() if bias is not None else None, M, N, K, x2.stride(0) if M > 1 else K, N,
                                      L.stream_ptr(xq.device)), "bq_gemm_packed_tn")
    return y.reshape(*xq.shape[:-1], N)


class _LinearBase(nn.Linear):
    def __init__(self, in_features: int, out_features: int, bias: bool = True, device=None, dtype=None,
                 config: dict = None) -> None:
        super().__init__(in_features, out_features, bias, device, dtype)
        self.config = config
        self.bypass = config.get("bypass", False)
        self.is_ptq = config.get("is_ptq", False)
        self.weight_requires_quantisation = True if self.is_ptq else False
        self.x_quantizer = None
        self.w_quantizer = None
        self.b_quantizer = None
        self._wq_bf16 = None          # derived bf16 cache of the quantised weight [N, K]
        self._wq_key = None
        self._wq_packed = None        # derived packed cache (PACKED_WEIGHTS): (key, packed uint8 [N, row_bytes], mismatching elements)
        if not self.bypass:
            self._setup_quantizers(config)

    # subclasses bind the three quantizers from config keys (reference linear.py:113-281)
    def _setup_quantizers(self, config: dict):
        raise NotImplementedError

    # ------------------------------------------------------------------ fused PTQ path
    def _fusable(self, x):
        if not x.is_cuda or x.dtype != torch.float32 or self.weight.dtype != torch.float32:
            return False
        if self.in_features % 8 != 0 or x.ndim not in (2, 3):
            return False
        try:
            kind, kw, _ = operand_format(self.config, "data_in")
            wkind, wkw, _ = operand_format(self.config, "weight")
        except KeyError:
            return False
        return significant_bits(kind, kw) <= 8 and significant_bits(wkind, wkw) <= 8

    def _weight_cache(self):
        w = self.weight
        key = (w.data_ptr(), w._version, w.device)
        if self._wq_bf16 is None or self._wq_key != key:
            # weight already holds quantised values (PTQ overwrite below); bf16 is an exact container for them
            canon = canonicalise(w.detach(), None, False, blocked=False)
            self._wq_bf16 = launch_quantize(w.detach(), make_format("none"), canon, out_dtype=torch.bfloat16)
            self._wq_key = key
        return self._wq_bf16

    def _packable(self):
        try:
            wkind, wkw, wbs = operand_format(self.config, "weight")
        except KeyError:
            return None
        if wkind != "block_fp" or not (2 <= wkw["width"] <= 8) or self.in_features % 256 or self.out_features % 32 or wbs is None:
            return None
        from ..quantizers.utils import resolve_block_shape

        if resolve_block_shape([self.out_features, self.in_features], wbs) != [1, 16]:
            return None
        return wkw

    def _packed_cache(self):
        """(packed weight, format kwargs) — w + 0.5 bits per element — or None when the weight format is not packable."""
        wkw = self._packable()
        if wkw is None:
            return None
        w = self.weight
        key = (w.data_ptr(), w._version, w.device)
        if self._wq_packed is None or self._wq_packed[0] != key:
            packed, bad = pack_weight(w.detach(), wkw["width"], wkw["exponent_width"], wkw["exponent_bias"])
            self._wq_packed = (key, packed, bad)
        return self._wq_packed[1], wkw

    def packed_bits_per_element(self):
        c = self._packed_cache()
        return None if c is None else 8.0 * c[0].numel() / (self.in_features * self.out_features)

    def _fused_forward(self, x):
        lib = L.load()
        kind, kw, block_size = operand_format(self.config, "data_in")
        K, N = self.in_features, self.out_features
        x2 = x.reshape(-1, K)
        if x2.stride(-1) != 1 or (x2.shape[0] > 1 and x2.stride(0) % 4 != 0):
            x2 = x2.contiguous()
        M = x2.shape[0]
        blocked = block_size is not None
        b0, b1 = 1, 1
        if blocked:
            cn = canonicalise(x.detach(), block_size, True, blocked=True)
            b0, b1 = cn.b0, cn.b1
        bias = self.bias.detach() if self.bias is not None else None
        if PACKED_WEIGHTS and M > 0 and b0 == 1:
            pc = self._packed_cache()
            if pc is not None:
                # x-quantizer (streaming kernel, bf16 out) -> GEMM that decodes the packed weights in its mainloop
                xq = quantize_operand_bf16(x2, kind, kw, [1, b1] if blocked else None, True)
                return gemm_packed(xq, pc[0], pc[1]["width"], pc[1]["exponent_width"], pc[1]["exponent_bias"], N, bias).reshape(
                    *x.shape[:-1], N)
        y = torch.empty((M, N), dtype=torch.float32, device=x.device)
        wq = self._weight_cache()
        if (FUSED_PROLOGUE and M > 0 and N > 0 and b0 == 1 and b1 == 16 and kind in ("block_fp", "block_minifloat")
                and K % 64 == 0 and N % 32 == 0):
            fmt = make_format(kind, b0=1, b1=16, fold=False, **kw)
            rc = lib.bq_linear_fused(ctypes.byref(fmt), x2.data_ptr(), M, K, x2.stride(0) if M > 1 else K, wq.data_ptr(), N,
                                     bias.data_ptr() if bias is not None else None, y.data_ptr(), N, L.stream_ptr(x.device))
            L.check(rc, "bq_linear_fused")
            return y.reshape(*x.shape[:-1], N)
        if M > 0 and N > 0:
            if b0 == 1:
                fmt = make_format(kind, b0=1, b1=b1, fold=False, **kw)
                nbytes = lib.bq_linear_workspace_bytes(ctypes.byref(fmt), M, K)
                ws = L.workspace(nbytes, x.device)
                rc = lib.bq_linear(ctypes.byref(fmt), x2.data_ptr(), M, K, x2.stride(0) if M > 1 else K, wq.data_ptr(), N,
                                   bias.data_ptr() if bias is not None else None, y.data_ptr(), N, ws.data_ptr(),
                                   ws.numel(), L.stream_ptr(x.device))
                L.check(rc, "bq_linear")
            else:
                # blocks spanning rows (e.g. data_in_block_size=[16] on a 3-D input): quantise with the general
                # kernel, then the same tensor-core GEMM
                xq = quantize_operand_bf16(x, kind, kw, block_size, True).reshape(M, K)
                rc = lib.bq_gemm_bf16_tn(xq.data_ptr(), wq.data_ptr(), y.data_ptr(),
                                         bias.data_ptr() if bias is not None else None, 1, M, N, K, K, K, N, 0, 0, 0,
                                         L.stream_ptr(x.device))
                L.check(rc, "bq_gemm_bf16_tn")
        return y.reshape(*x.shape[:-1], N)

    def _ensure_ptq(self):
        """One-off PTQ overwrite of weight / bias with their quantised values (reference linear.py:66-70)."""
        if self.weight_requires_quantisation:
            with torch.no_grad():
                self.weight.copy_(self.w_quantizer(self.weight.data))
                if self.bias is not None:
                    self.bias.copy_(self.b_quantizer(self.bias.data))
            self.weight_requires_quantisation = False
            self._wq_bf16 = None
            self._wq_packed = None

    def accepts_prequantized(self) -> bool:
        """True when `forward_prequantized` may be used: PTQ mode and a weight format that is exact in bf16."""
        if self.bypass or not self.is_ptq or self.weight.dtype != torch.float32 or self.in_features % 8 != 0:
            return False
        try:
            wkind, wkw, _ = operand_format(self.config, "weight")
        except KeyError:
            return False
        return significant_bits(wkind, wkw) <= 8

    @torch.no_grad()
    def forward_prequantized(self, xq: torch.Tensor, *, scale: float = 1.0, relu: bool = False, residual: torch.Tensor = None,
                             out_format=None, out_blocks_along_rows: bool = False, out: torch.Tensor = None,
                             peer_out_ptrs=None) -> torch.Tensor:
        """
        y = F.linear(xq, Wq, bq) for an input that ALREADY went through this module's x-quantizer inside the kernel that
        produced it (bf16 carrier of the exact quantised values), with the layer glue that follows fused into the GEMM
        epilogue (bq_gemm_bf16_tn_ex, include/bq.h), in the reference's op order:
            y = y * scale;  y = relu(y);  y = residual + y;  y = Q_out(y)
        out_format: None -> fp32 result; (kind, kwargs) of the x-quantizer of the NEXT op (block [1,16]) -> bf16 result
        holding its exact quantised values.  out_blocks_along_rows: the 16-blocks run over 16 consecutive rows (tokens)
        instead of 16 consecutive features (the k^T operand of bmm_0).
        out: optional preallocated [M, N] destination (row stride >= N: e.g. this rank's column slab of a gathered
        [M, N_total] buffer).  peer_out_ptrs: device addresses of the SAME slab in up to 7 peer-mapped buffers with the same
        row stride — the epilogue stores every tile there too (fused all-gather of the column-parallel Linear, dist.py).
        """
        assert xq.dtype == torch.bfloat16 and xq.is_cuda and xq.shape[-1] == self.in_features
        self._ensure_ptq()
        lib = L.load()
        K, N = self.in_features, self.out_features
        x2 = xq.reshape(-1, K)
        if x2.stride(-1) != 1 or (x2.shape[0] > 1 and x2.stride(0) % 8 != 0):
            x2 = x2.contiguous()
        M = x2.shape[0]
        out_dtype = torch.float32 if out_format is None else torch.bfloat16
        if out is None:
            y = torch.empty((M, N), dtype=out_dtype, device=xq.device)
        else:
            y = out
            if tuple(y.shape) != (M, N) or y.dtype != out_dtype or y.device != xq.device or (N > 1 and y.stride(1) != 1):
                raise ValueError(f"out must be a [{M}, {N}] {out_dtype} tensor with unit column stride on {xq.device}")
        ldc = y.stride(0) if M > 1 else max(N, y.stride(0))
        peers = list(peer_out_ptrs or [])
        if M > 0 and N > 0:
            wq = self._weight_cache()
            bias = self.bias.detach() if self.bias is not None else None
            lda = x2.stride(0) if M > 1 else K
            plain = scale == 1.0 and not relu and residual is None and out_format is None and not peers
            if plain:
                rc = lib.bq_gemm_bf16_tn(x2.data_ptr(), wq.data_ptr(), y.data_ptr(), bias.data_ptr() if bias is not None else None,
                                         1, M, N, K, lda, K, ldc, 0, 0, 0, L.stream_ptr(xq.device))
                L.check(rc, "bq_gemm_bf16_tn")
            else:
                ep = L.BqGemmEpilogue()
                ep.bias = bias.data_ptr() if bias is not None else None
                res2 = None
                if residual is not None:
                    res2 = residual.reshape(M, N)
                    if res2.stride(-1) != 1 or res2.dtype != torch.float32:
                        res2 = res2.float().contiguous()
                    ep.residual, ep.ldr = res2.data_ptr(), res2.stride(0) if M > 1 else N
                ep.scale, ep.act = float(scale), 1 if relu else 0
                ep.out_dtype = L.BQ_F32 if out_format is None else L.BQ_BF16
                fmt = None
                if out_format is not None:
                    kind, kw = out_format
                    fmt = make_format(kind, b0=1, b1=16, **kw)
                    ep.qfmt = ctypes.pointer(fmt)
                    ep.qdir = 1 if out_blocks_along_rows else 0
                if len(peers) > 7:
                    raise ValueError("at most 7 peer replicas (8 GPUs per NVSwitch domain)")
                ep.n_replicas = len(peers)
                for i, ptr in enumerate(peers):
                    ep.replicas[i] = int(ptr)
                rc = lib.bq_gemm_bf16_tn_ex(x2.data_ptr(), wq.data_ptr(), y.data_ptr(), ctypes.byref(ep), M, N, K, lda, K, ldc,
                                            L.stream_ptr(xq.device))
                L.check(rc, "bq_gemm_bf16_tn_ex")
        return y if out is not None else y.reshape(*xq.shape[:-1], N)

    def forward(self, x):
        if self.bypass:
            return F.linear(x, self.weight, self.bias)
        elif self.is_ptq:
            # the reference quantises under no_grad but runs F.linear OUTSIDE it (linear.py:63-71): weight / bias receive
            # gradients when autograd is on.  The fused kernels are forward-only, so they serve only the no-grad case.
            wants_grad = torch.is_grad_enabled() and (self.weight.requires_grad
                                                      or (self.bias is not None and self.bias.requires_grad))
            with torch.no_grad():
                self._ensure_ptq()
                if not wants_grad and self._fusable(x):
                    return self._fused_forward(x)
                x = self.x_quantizer(x)
            return F.linear(x, self.weight, self.bias)
        else:
            x = self.x_quantizer(x)
            w = self.w_quantizer(self.weight)
            bias = self.b_quantizer(self.bias) if self.bias is not None else None
            return F.linear(x, w, bias)

    @classmethod
    def from_float(cls, linear_fp32: nn.Linear, config: dict):
        linear = cls(linear_fp32.in_features, linear_fp32.out_features, bias=linear_fp32.bias is not None, config=config)
        with torch.no_grad():
            linear.weight.copy_(linear_fp32.weight)
            if linear.bias is not None:
                linear.bias.copy_(linear_fp32.bias)
        return linear.to(linear_fp32.weight.device)

    def __repr__(self):
        return "{}(in_features={}, out_features={}, bias={}, bypass={}, is_ptq={}, x/w/b-width={}/{}/{})".format(
            self.__class__.__name__, self.in_features, self.out_features, self.bias is not None, self.bypass, self.is_ptq,
            self.config.get("data_in_width", "NA"), self.config.get("weight_width", "NA"), self.config.get("bias_width", "NA"))


ROPE_EPILOGUE = True      # False: q / k GEMMs with fp32 outputs + the RoPE+quantise kernels (A/B, tests)


def rope_epilogue_fusable(lin: "_LinearBase", head_dim: int) -> bool:
    return (ROPE_EPILOGUE and lin.accepts_prequantized() and head_dim in (64, 128) and lin.out_features % head_dim == 0
            and (lin.bias is None or lin.bias.data_ptr() % 16 == 0))


@torch.no_grad()
def rope_prequantized(lin: "_LinearBase", xq: torch.Tensor, cos_t: torch.Tensor, sin_t: torch.Tensor, pos, fmt, seq_len: int, head_dim: int,
                      along_rows: bool) -> torch.Tensor:
    """Q_fmt(rope(lin(xq))) as bf16 [rows, N] in ONE GEMM launch (bq_gemm_bf16_tn_rope): q_proj / k_proj of a Llama layer with the
    rotary embedding and matmul_0's operand quantizer in the epilogue (reference modeling_llama.py:274-276, :309-314).  `fmt`: the
    bq_format of the operand (rope_quantize_operands), along_rows: blocks of 16 consecutive tokens (the k^T operand)."""
    assert xq.dtype == torch.bfloat16 and xq.is_cuda and xq.shape[-1] == lin.in_features
    lin._ensure_ptq()
    lib = L.load()
    K, N = lin.in_features, lin.out_features
    x2 = xq.reshape(-1, K)
    if x2.stride(-1) != 1 or (x2.shape[0] > 1 and x2.stride(0) % 8 != 0):
        x2 = x2.contiguous()
    M = x2.shape[0]
    y = torch.empty((M, N), dtype=torch.bfloat16, device=xq.device)
    if M > 0:
        wq = lin._weight_cache()
        bias = lin.bias.detach() if lin.bias is not None else None
        rc = lib.bq_gemm_bf16_tn_rope(x2.data_ptr(), wq.data_ptr(), y.data_ptr(), bias.data_ptr() if bias is not None else None,
                                      ctypes.byref(fmt), 1 if along_rows else 0, cos_t.data_ptr(), sin_t.data_ptr(),
                                      pos.data_ptr() if pos is not None else None, cos_t.shape[0], seq_len, head_dim, M, N, K,
                                      x2.stride(0) if M > 1 else K, K, N, L.stream_ptr(xq.device))
        L.check(rc, "bq_gemm_bf16_tn_rope")
    return y


QKV_ONE_LAUNCH = True     # False: three GEMM launches for q / k / v (A/B, tests)


def qkv_rope_fusable(q: "_LinearBase", k: "_LinearBase", v: "_LinearBase", head_dim: int) -> bool:
    return (QKV_ONE_LAUNCH and all(rope_epilogue_fusable(m, head_dim) for m in (q, k)) and v.accepts_prequantized()
            and q.in_features == k.in_features == v.in_features and q.out_features == k.out_features == v.out_features
            and q.out_features % 256 == 0 and len({m.bias is None for m in (q, k, v)}) == 1)


@torch.no_grad()
def qkv_rope_prequantized(q: "_LinearBase", k: "_LinearBase", v: "_LinearBase", xq: torch.Tensor, cos_t, sin_t, pos, fq, fk, v_format,
                          seq_len: int, head_dim: int):
    """(Q_q(rope(q_proj(x))), Q_k(rope(k_proj(x))) along tokens, Q_v(v_proj(x))) as bf16 [rows, H] each, in ONE GEMM launch over the three
    quantised weights concatenated along N (bq_gemm_bf16_tn_qkv_rope) — for a Llama layer whose q / k / v projections share their
    x-quantizer (reference modeling_llama.py:274-276, :309-314, :341-344).  Same bits as three separate launches."""
    assert xq.dtype == torch.bfloat16 and xq.is_cuda and xq.shape[-1] == q.in_features
    for m in (q, k, v):
        m._ensure_ptq()
    lib = L.load()
    K, H = q.in_features, q.out_features
    wcat, bcat = _qkv_concat_cache(q, k, v)
    x2 = xq.reshape(-1, K)
    if x2.stride(-1) != 1 or (x2.shape[0] > 1 and x2.stride(0) % 8 != 0):
        x2 = x2.contiguous()
    M = x2.shape[0]
    outs = [torch.empty((M, H), dtype=torch.bfloat16, device=xq.device) for _ in range(3)]
    if M > 0:
        vk, vkw = v_format
        fv = make_format(vk, b0=1, b1=16, **vkw)
        rc = lib.bq_gemm_bf16_tn_qkv_rope(x2.data_ptr(), wcat.data_ptr(), outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(),
                                          bcat.data_ptr() if bcat is not None else None, ctypes.byref(fq), ctypes.byref(fk), ctypes.byref(fv),
                                          cos_t.data_ptr(), sin_t.data_ptr(), pos.data_ptr() if pos is not None else None, cos_t.shape[0],
                                          seq_len, head_dim, M, H, K, x2.stride(0) if M > 1 else K, K, H, L.stream_ptr(xq.device))
        L.check(rc, "bq_gemm_bf16_tn_qkv_rope")
    return outs


def _qkv_concat_cache(q: "_LinearBase", k: "_LinearBase", v: "_LinearBase"):
    """(weights [3H, K] bf16, bias [3H] fp32 or None) of three projections that read the same operand, built once per weight version."""
    key = tuple((m.weight.data_ptr(), m.weight._version) for m in (q, k, v)) + (q.weight.device,)
    cache = getattr(q, "_qkv_cache", None)
    if cache is None or cache[0] != key:
        wcat = torch.cat([m._weight_cache() for m in (q, k, v)], dim=0).contiguous()
        bcat = torch.cat([m.bias.detach() for m in (q, k, v)]).contiguous() if q.bias is not None else None
        q._qkv_cache = cache = (key, wcat, bcat)
        q._wq_bf16 = k._wq_bf16 = v._wq_bf16 = None          # the separate bf16 copies are rebuilt on demand
    return cache[1], cache[2]


def qkv_plain_fusable(q: "_LinearBase", k: "_LinearBase", v: "_LinearBase") -> bool:
    return (QKV_ONE_LAUNCH and all(m.accepts_prequantized() for m in (q, k, v)) and q.in_features == k.in_features == v.in_features
            and q.out_features == k.out_features == v.out_features and len({m.bias is None for m in (q, k, v)}) == 1)


@torch.no_grad()
def qkv_plain_prequantized(q: "_LinearBase", k: "_LinearBase", v: "_LinearBase", xq: torch.Tensor):
    """(q_proj(x), k_proj(x), v_proj(x)) as fp32 column views [rows, H] (row stride 3H) of ONE GEMM over the concatenated quantised
    weights — for layers whose attention keeps fp32 operands (block_log split path) and whose three projections share an x-quantizer.
    Same bits as three launches (an output column's K reduction does not depend on its neighbours)."""
    assert xq.dtype == torch.bfloat16 and xq.is_cuda and xq.shape[-1] == q.in_features
    for m in (q, k, v):
        m._ensure_ptq()
    lib = L.load()
    K, H = q.in_features, q.out_features
    wcat, bcat = _qkv_concat_cache(q, k, v)
    x2 = xq.reshape(-1, K)
    if x2.stride(-1) != 1 or (x2.shape[0] > 1 and x2.stride(0) % 8 != 0):
        x2 = x2.contiguous()
    M = x2.shape[0]
    y = torch.empty((M, 3 * H), dtype=torch.float32, device=xq.device)
    if M > 0:
        rc = lib.bq_gemm_bf16_tn(x2.data_ptr(), wcat.data_ptr(), y.data_ptr(), bcat.data_ptr() if bcat is not None else None, 1, M, 3 * H, K,
                                 x2.stride(0) if M > 1 else K, K, 3 * H, 0, 0, 0, L.stream_ptr(xq.device))
        L.check(rc, "bq_gemm_bf16_tn(q|k|v)")
    return y[:, :H], y[:, H:2 * H], y[:, 2 * H:]


GATED_EPILOGUE = True     # False: gate / up GEMMs + the silu*mul quantizer kernel (A/B, tests)


def gated_silu_fusable(gate: "_LinearBase", up: "_LinearBase") -> bool:
    """Whether `gated_silu_prequantized` can serve this gate / up pair: both PTQ with bf16-exact weights, no bias, the same shape,
    whole blocks of 16 features."""
    return (GATED_EPILOGUE and gate.accepts_prequantized() and up.accepts_prequantized() and gate.bias is None and up.bias is None
            and gate.in_features == up.in_features and gate.out_features == up.out_features and gate.out_features % 16 == 0
            and gate.out_features >= 64)


@torch.no_grad()
def gated_silu_prequantized(gate: "_LinearBase", up: "_LinearBase", xq: torch.Tensor, out_format) -> torch.Tensor:
    """Q_out(silu(gate(xq)) * up(xq)) in ONE GEMM launch — the operand of Llama's down_proj (reference
    models/llama_quantized/modeling_llama.py:84 `down_proj(act_fn(gate_proj(x)) * up_proj(x))`, x-quantizer of
    quantized_modules/linear.py:63-71) for an input that already went through the (identical) x-quantizers of gate_proj and up_proj.
    The two quantised weights are interleaved in groups of 16 rows ([gate f..f+16), [up f..f+16), ...) so that a 32-column chunk of
    the accumulator holds one block of the consumer's x-quantizer; the epilogue (bq_gemm_bf16_tn_ex, act = 2) applies silu * up and
    the quantizer and stores bf16 [M, F].  Bit-identical to gate GEMM + up GEMM + silu_mul_quantize (same accumulation order per
    output column, same element-wise code), without the 16 B/element fp32 round trip between them."""
    assert xq.dtype == torch.bfloat16 and xq.is_cuda and xq.shape[-1] == gate.in_features
    gate._ensure_ptq()
    up._ensure_ptq()
    lib = L.load()
    K, Fo = gate.in_features, gate.out_features
    key = (gate.weight.data_ptr(), gate.weight._version, up.weight.data_ptr(), up.weight._version, gate.weight.device)
    cache = getattr(gate, "_gu_cache", None)
    if cache is None or cache[0] != key:
        wg, wu = gate._weight_cache(), up._weight_cache()
        wgu = torch.stack((wg.view(Fo // 16, 16, K), wu.view(Fo // 16, 16, K)), dim=1).reshape(2 * Fo, K).contiguous()
        gate._gu_cache = cache = (key, wgu)
        gate._wq_bf16 = up._wq_bf16 = None      # the separate bf16 copies are rebuilt on demand (op-by-op path); no need to hold both
    wgu = cache[1]
    x2 = xq.reshape(-1, K)
    if x2.stride(-1) != 1 or (x2.shape[0] > 1 and x2.stride(0) % 8 != 0):
        x2 = x2.contiguous()
    M = x2.shape[0]
    y = torch.empty((M, Fo), dtype=torch.bfloat16, device=xq.device)
    if M > 0:
        kind, kw = out_format
        fmt = make_format(kind, b0=1, b1=16, **kw)
        ep = L.BqGemmEpilogue()
        ep.scale, ep.act, ep.out_dtype = 1.0, 2, L.BQ_BF16
        ep.qfmt = ctypes.pointer(fmt)
        ep.qdir = 0
        rc = lib.bq_gemm_bf16_tn_ex(x2.data_ptr(), wgu.data_ptr(), y.data_ptr(), ctypes.byref(ep), M, 2 * Fo, K,
                                    x2.stride(0) if M > 1 else K, K, Fo, L.stream_ptr(xq.device))
        L.check(rc, "bq_gemm_bf16_tn_ex(gated silu)")
    return y.reshape(*xq.shape[:-1], Fo)


def _bind(self, quantizer, keys, config, with_blocks):
    """Bind x / w / b quantizers from `<prefix>_<key>` entries (x: skip_first_dim=True, w/b: False)."""
    def one(prefix, skip):
        kw = {k: config[f"{prefix}_{k}"] for k in keys}
        if with_blocks:
            kw["block_size"] = config[f"{prefix}_block_size"]
            kw["skip_first_dim"] = skip
        return partial(quantizer, **kw)

    self.x_quantizer = one("data_in", True)
    self.w_quantizer = one("weight", False)
    self.b_quantizer = one("bias", False) if self.bias is not None else None


class LinearBlockFP(_LinearBase):
    def _setup_quantizers(self, config: dict):
        _bind(self, block_fp_quantizer, ("width", "exponent_width", "exponent_bias"), config, True)


class LinearBlockMinifloat(_LinearBase):
    def _setup_quantizers(self, config: dict):
        _bind(self, block_minifloat_quantizer, ("width", "exponent_width", "exponent_bias_width"), config, True)


class LinearBlockLog(_LinearBase):
    def _setup_quantizers(self, config: dict):
        _bind(self, block_log_quantizer, ("width", "exponent_bias_width"), config, True)


class LinearMinifloatDenorm(_LinearBase):
    def _setup_quantizers(self, config: dict):
        _bind(self, minifloat_denorm_quantizer, ("width", "exponent_width", "exponent_bias"), config, False)


class LinearMinifloatIEEE(_LinearBase):
    def _setup_quantizers(self, config: dict):
        _bind(self, minifloat_ieee_quantizer, ("width", "exponent_width", "exponent_bias"), config, False)


class LinearInteger(_LinearBase):
    def _setup_quantizers(self, config: dict):
        def one(prefix):
            return partial(integer_quantizer, width=config[f"{prefix}_width"], frac_width=config[f"{prefix}_frac_width"],
                           is_signed=True)

        self.x_quantizer = one("data_in")
        self.w_quantizer = one("weight")
        self.b_quantizer = one("bias") if self.bias is not None else None
