"""
Fused causal quantized attention (host side of bq_attention_causal, include/bq.h).

No single reference function corresponds to this: it is the composition the reference's attention modules
spell out op by op — bmm_0 / matmul_0, causal mask + max(finfo.min), softmax, bmm_1 / matmul_1
(models/opt_quantized/modeling_opt.py:246-312, models/llama_quantized/modeling_llama.py:309-344) — executed
by one kernel that keeps scores and probabilities on chip.  The quantized model classes call it when the
mask is purely causal and the layer's two matmul configs are block formats the kernel supports; otherwise
they fall back to the op-by-op path through QUANTIZED_FUNC_MAP.
"""
from __future__ import annotations

import ctypes

import torch

from .... import _lib as L
from ..quantized_modules.linear import operand_format, quantize_operand_bf16, significant_bits
from ..quantizers.utils import make_format, resolve_block_shape

_SUPPORTED_P = ("block_fp", "block_minifloat")


def fusable(cfg0: dict, cfg1: dict, head_dim: int, seq_len: int) -> bool:
    """Can (bmm_0 cfg, bmm_1 cfg) run on the fused kernel?"""
    try:
        if cfg0.get("bypass", False) or cfg1.get("bypass", False):
            return False
        if head_dim not in (64, 128):
            return False
        (qk, qkw, qbs), (kk, kkw, kbs) = operand_format(cfg0, "data_in"), operand_format(cfg0, "weight")
        (pk, pkw, pbs), (vk, vkw, vbs) = operand_format(cfg1, "data_in"), operand_format(cfg1, "weight")
    except KeyError:
        return False
    if pk not in _SUPPORTED_P or kk not in ("block_fp", "block_minifloat") or any(b is None for b in (qbs, kbs, pbs, vbs)):
        return False
    if max(significant_bits(qk, qkw), significant_bits(kk, kkw), significant_bits(pk, pkw), significant_bits(vk, vkw)) > 8:
        return False
    if qk == "block_log" or vk == "block_log":
        return False                       # tensor-global zero-block rule: keep the op-by-op path
    # blocks the reference would infer on the flattened [B*h, rows, cols] operands
    qb = resolve_block_shape([1, seq_len, head_dim], qbs)
    kb = resolve_block_shape([1, head_dim, seq_len], kbs)
    pb = resolve_block_shape([1, seq_len, seq_len], pbs)
    vb = resolve_block_shape([1, seq_len, head_dim], vbs)
    ok_last = lambda b, n: b[1] == 1 and b[2] in (4, 8, 16, 32, 64) and n % b[2] == 0
    return ok_last(qb, head_dim) and ok_last(vb, head_dim) and kb[1] == 1 and kb[2] in (1, 2, 4, 8, 16, 32, 64) and \
        pb[1] == 1 and pb[2] == 16


def output_quantizable(out_cfg: dict, hidden: int, rows: int = 1) -> bool:
    """Can the x-quantizer of the Linear that consumes the attention output run in the attention epilogue?
    `rows` = S: the reference feeds that Linear the 3-D [B, S, H] attention output, so a short block size such as [16]
    resolves to [1, S, 16] (quantizers/utils.py:42-67) and is NOT a row block."""
    try:
        if out_cfg is None or out_cfg.get("bypass", False) or not out_cfg.get("is_ptq", False):
            return False
        ok, okw, obs = operand_format(out_cfg, "data_in")
    except KeyError:
        return False
    if ok not in ("block_fp", "block_minifloat") or obs is None or significant_bits(ok, okw) > 8:
        return False
    ob = resolve_block_shape([1, max(int(rows), 1), hidden], obs)
    return ob[1] == 1 and ob[2] == 16


def quantize_qkv(q, k, v, cfg0: dict, cfg1: dict, num_heads: int):
    """fp32 [B, S, H] projections -> the three bf16 [B, S, H] operands of the fused kernel."""
    B, S, H = q.shape
    d = H // num_heads
    (qk, qkw, qbs), (kk, kkw, kbs) = operand_format(cfg0, "data_in"), operand_format(cfg0, "weight")
    (vk, vkw, vbs) = operand_format(cfg1, "weight")
    # q / v: blocks along d — identical to blocking the [B*S, H] matrix along H because b1 divides d
    qb = resolve_block_shape([1, S, d], qbs)[2]
    vb = resolve_block_shape([1, S, d], vbs)[2]
    Qq = quantize_operand_bf16(q.reshape(B * S, H), qk, qkw, [1, qb], True)
    Vq = quantize_operand_bf16(v.reshape(B * S, H), vk, vkw, [1, vb], True) if v is not None else None
    # k: the reference quantises k^T, i.e. blocks of consecutive KEY POSITIONS at fixed feature
    kb = resolve_block_shape([1, d, S], kbs)[2]
    Kq = quantize_operand_bf16(k.transpose(1, 2), kk, kkw, [1, kb], True, transpose_out=True)     # -> [B, S, H]
    return Qq, Kq, Vq


def key_mask_bits(valid: torch.Tensor) -> torch.Tensor:
    """[B, S] boolean / 0-1 key-validity mask -> int32 bitmap [B, 4 * ceil(S / 128)] for bq_attention_masked: bit i of word w set
    = key 32 * w + i takes part; keys >= S are cleared (the kernel's key tiles are zero-filled there)."""
    B, S = valid.shape
    words = 4 * ((S + 127) // 128)
    v = torch.zeros((B, words * 32), dtype=torch.int64, device=valid.device)
    v[:, :S] = valid.to(torch.int64)
    w = (v.view(B, words, 32) << torch.arange(32, device=valid.device, dtype=torch.int64)).sum(-1)      # < 2^32
    return (w & 0xFFFFFFFF).to(torch.int64).where(w < 2 ** 31, w - 2 ** 32).to(torch.int32).contiguous()


def causal_key_mask(attention_mask: torch.Tensor):
    """Key-padding bitmap for a decoder batch, or None when the fused kernel must not be used: with the causal mask every query
    row needs one visible key, which holds for every row iff key 0 of every sequence is valid (right padding).  Fully masked rows
    (left padding) are a uniform distribution over ALL keys in the reference (finfo.min everywhere after the clamp,
    opt_quantized/modeling_opt.py:520-548, :266-270) — the op-by-op path reproduces that.  One device-to-host look."""
    if attention_mask is None or attention_mask.ndim != 2:
        return None
    valid = attention_mask != 0
    if not bool(valid[:, 0].all()):
        return None
    return key_mask_bits(valid)


def fused_causal_attention_q(Qq: torch.Tensor, Kq: torch.Tensor, Vq: torch.Tensor, cfg1: dict, num_heads: int, B: int, S: int,
                             score_div: float = 1.0, out_cfg: dict = None, out: torch.Tensor = None, causal: bool = True,
                             key_mask: torch.Tensor = None) -> torch.Tensor:
    """Kernel call on already-quantised bf16 operands ([B*S, H] or [B, S, H], unit stride along H).
    Returns fp32 [B, S, H], or — with `out_cfg` (config of the Linear consuming the result) — its bf16 x-quantised form.
    `out` (with `out_cfg`): preallocated bf16 [B*S, H] destination with unit column stride and any row stride — e.g. this rank's
    column slab of a gathered [B*S, H_total] buffer (tensor-parallel layer, dist.py)."""
    lib = L.load()
    H = Qq.shape[-1]
    d = H // num_heads
    (pk, pkw, pbs) = operand_format(cfg1, "data_in")
    fp = make_format(pk, b0=1, b1=16, **pkw)
    dev = Qq.device
    ld = lambda t: t.stride(-2)
    if not causal or key_mask is not None:
        # bidirectional (BERT) and / or key-padding mask: `key_mask` = key_mask_bits(valid keys); built here when absent
        if key_mask is None:
            key_mask = key_mask_bits(torch.ones((B, S), dtype=torch.bool, device=dev))
        if key_mask.dtype != torch.int32 or key_mask.device != dev or not key_mask.is_contiguous() or key_mask.shape[0] != B:
            raise ValueError("key_mask must be the int32 [B, words] bitmap of key_mask_bits() on the operands' device")
        fo = None
        if out_cfg is not None:
            ok, okw, _ = operand_format(out_cfg, "data_in")
            fo = make_format(ok, b0=1, b1=16, **okw)
        ldo = H
        if out is None:
            out = torch.empty((B, S, H), dtype=torch.float32 if fo is None else torch.bfloat16, device=dev)
        else:
            ldo = out.stride(-2)
        rc = lib.bq_attention_masked(ctypes.byref(fp), ctypes.byref(fo) if fo is not None else None, Qq.data_ptr(), Kq.data_ptr(),
                                     Vq.data_ptr(), out.data_ptr(), B, num_heads, S, d, ld(Qq), ld(Kq), ld(Vq), ldo, float(score_div),
                                     1 if causal else 0, key_mask.data_ptr(), key_mask.shape[1], L.stream_ptr(dev))
        L.check(rc, "bq_attention_masked")
        return out
    if out_cfg is None:
        out = torch.empty((B, S, H), dtype=torch.float32, device=dev)
        rc = lib.bq_attention_causal(ctypes.byref(fp), Qq.data_ptr(), Kq.data_ptr(), Vq.data_ptr(), out.data_ptr(), B, num_heads,
                                     S, d, ld(Qq), ld(Kq), ld(Vq), H, float(score_div), L.stream_ptr(dev))
        L.check(rc, "bq_attention_causal")
        return out
    ok, okw, _ = operand_format(out_cfg, "data_in")
    fo = make_format(ok, b0=1, b1=16, **okw)
    ldo = H
    if out is None:
        out = torch.empty((B, S, H), dtype=torch.bfloat16, device=dev)
    else:
        if out.dtype != torch.bfloat16 or out.device != dev or out.shape[-1] != H or out.stride(-1) != 1 or out.numel() != B * S * H:
            raise ValueError(f"out must be a bf16 [{B * S}, {H}] tensor with unit column stride on {dev}")
        ldo = out.stride(-2)
    rc = lib.bq_attention_causal_q(ctypes.byref(fp), ctypes.byref(fo), Qq.data_ptr(), Kq.data_ptr(), Vq.data_ptr(), out.data_ptr(),
                                   B, num_heads, S, d, ld(Qq), ld(Kq), ld(Vq), ldo, float(score_div), L.stream_ptr(dev))
    L.check(rc, "bq_attention_causal_q")
    return out


def fused_causal_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, cfg0: dict, cfg1: dict, num_heads: int,
                           score_div: float = 1.0, out_cfg: dict = None, causal: bool = True,
                           key_mask: torch.Tensor = None) -> torch.Tensor:
    """
    q, k, v: fp32 [B, S, H] projections (q already scaled where the model scales before bmm_0, as OPT does).
    Returns fp32 [B, S, H] = concat over heads of  Q(softmax(Q(q) Q(k)^T / score_div, causal)) @ Q(v)
    (bf16, x-quantised for the consuming Linear, when `out_cfg` is given).
    """
    B, S, H = q.shape
    Qq, Kq, Vq = quantize_qkv(q, k, v, cfg0, cfg1, num_heads)
    return fused_causal_attention_q(Qq, Kq, Vq, cfg1, num_heads, B, S, score_div, out_cfg, causal=causal, key_mask=key_mask)
