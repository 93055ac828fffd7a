"""
Attention for matmul configs that keep an UNQUANTISED fp32 operand — host side of bq_rope_quantize_split, bq_split3_bf16_transposed,
bq_bmm_split_tn and bq_softmax_quantize (include/bq.h).

The reference's block_log matmul quantises x only and multiplies by the fp32 y as is (quantized_functions/matmul.py:286-297), so
the one-kernel attention (attention.py), whose operands are all bf16-exact, does not apply.  The op-by-op path for that case cost
3.7x the block_minifloat model (profiles/r01_profile_llama_block_log_s9.txt: a third of the step in torch's mask-add / clamp /
softmax passes over the B*h*S*S scores, the rest in quantizer launches on S x S tensors).  Here, per attention call:

    q, k (fp32, token-major)  --bq_rope_quantize_split-->  Qq bf16 [B,h,S,d] (matmul_0's x-quantizer)  +  k = k0+k1+k2 bf16 planes
    v                         --bq_split3_bf16_transposed-> planes of v^T [B,3,h,d,S]
    scores = Qq @ (k0+k1+k2)^T                               bq_bmm_split_tn, 3 exact plane products, fp32 accumulate (TMEM)
    P = Q_x(softmax(max(scores / div + mask, finfo.min)))    bq_softmax_quantize: one read of the scores, one bf16 write
    out = P @ (v0+v1+v2)                                     bq_bmm_split_tn, written token-major [B,S,H]

Products of a power of two with a bf16 plane are exact, so both matmuls equal the reference's fp32 matmuls up to accumulation order.
block_log values travel in bf16 under the carrier rule of include/bq.h (outputs below 2^-126 become 0 or 2^-126).
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from .... import _lib as L
from ..quantized_modules.linear import operand_format
from ..quantizers.utils import make_format, resolve_block_shape
from .rotary_positional_encoding import _table_quantizer

_X_KINDS = ("block_log", "block_fp", "block_minifloat")


def _x_format(cfg: dict, shape3):
    """(kind, kwargs) of the x operand when its blocks resolve to [1, 16] along the last dim of the flattened 3-D operand."""
    kind, kw, bs = operand_format(cfg, "data_in")
    if kind not in _X_KINDS or bs is None:
        return None
    if resolve_block_shape(list(shape3), bs)[1:] != [1, 16]:
        return None
    if kind != "block_log":
        from ..quantized_modules.linear import significant_bits
        if significant_bits(kind, kw) > 8:
            return None
    return kind, kw


def splittable(cfg0: dict, cfg1: dict, head_dim: int, seq_len: int) -> bool:
    """Do (matmul_0 cfg, matmul_1 cfg) take the split path?  Both must leave y unquantised (block_log / log) and block x in 16s."""
    try:
        if cfg0.get("bypass", False) or cfg1.get("bypass", False):
            return False
        if cfg0["name"] not in ("block_log", "log") or cfg1["name"] not in ("block_log", "log"):
            return False
        if head_dim % 32 or seq_len % 16:
            return False
        return _x_format(cfg0, [1, seq_len, head_dim]) is not None and _x_format(cfg1, [1, seq_len, seq_len]) is not None
    except KeyError:
        return False


def rope_quantize_split(q: torch.Tensor, k: torch.Tensor, cos, sin, position_ids, rope_cfg: Optional[dict], cfg0: dict, num_heads: int):
    """q, k fp32 [B, S, H] -> (Qq bf16 [B, h, S, d], K planes bf16 [B, 3, h, S, d]).  `rope_cfg` None: no rotation (OPT-style
    attention); else the Llama rotary embedding with the tables quantised as the reference does (bq_rope_quantize_split)."""
    B, S, H = q.shape
    d = H // num_heads
    kind, kw = _x_format(cfg0, [1, S, d])
    fq = make_format(kind, b0=1, b1=16, **kw)
    pos = None
    cos_p = sin_p = None
    rows = S
    if rope_cfg is not None:
        tq = _table_quantizer(rope_cfg, rope_cfg["name"])
        cos_t = tq(cos.squeeze(1).squeeze(0)).contiguous()
        sin_t = tq(sin.squeeze(1).squeeze(0)).contiguous()
        if cos_t.dtype != torch.float32 or cos_t.shape[-1] != d:
            raise NotImplementedError("rotary tables must quantise to fp32 [rows, head_dim]")
        rows = cos_t.shape[0]
        if position_ids is not None:
            pos = position_ids.expand(B, S).to(torch.int64)
            lo, hi = (int(v) for v in torch.stack((pos.min(), pos.max())).tolist())
            if lo < -rows or hi >= rows:
                raise IndexError(f"index {hi if hi >= rows else lo} is out of bounds for dimension 0 with size {rows}")
            if lo < 0:
                pos = torch.where(pos < 0, pos + rows, pos)
            pos = pos.contiguous()
        cos_p, sin_p = cos_t.data_ptr(), sin_t.data_ptr()
    qc = q if (q.stride(-1) == 1 and q.stride(0) == S * q.stride(1)) else q.contiguous()
    kc = k if (k.stride(-1) == 1 and k.stride(0) == S * k.stride(1)) else k.contiguous()
    Qq = torch.empty((B, num_heads, S, d), dtype=torch.bfloat16, device=q.device)
    Kp = torch.empty((B, 3, num_heads, S, d), dtype=torch.bfloat16, device=q.device)
    rc = L.load().bq_rope_quantize_split(qc.data_ptr(), kc.data_ptr(), cos_p, sin_p, pos.data_ptr() if pos is not None else None, rows,
                                         B, S, num_heads, d, qc.stride(1), kc.stride(1), ctypes.byref(fq), Qq.data_ptr(), Kp.data_ptr(),
                                         L.stream_ptr(q.device))
    L.check(rc, "bq_rope_quantize_split")
    return Qq, Kp


_TERMS3 = ((0, 2), (0, 1), (0, 0))                        # smallest plane first
_TA = (ctypes.c_int32 * 3)(*[t[0] for t in _TERMS3])
_TB = (ctypes.c_int32 * 3)(*[t[1] for t in _TERMS3])


def split_attention(Qq: torch.Tensor, Kp: torch.Tensor, v: torch.Tensor, cfg1: dict, num_heads: int, score_div: float = 1.0,
                    causal: bool = True, key_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Qq bf16 [B, h, S, d], K planes bf16 [B, 3, h, S, d] (rope_quantize_split), v fp32 [B, S, H] -> fp32 [B, S, H]
    = concat over heads of  Q_x(softmax(Qq k^T / score_div + mask)) @ v   with k, v at full fp32 precision."""
    lib = L.load()
    B, h, S, d = Qq.shape
    H = h * d
    dev = Qq.device
    st = L.stream_ptr(dev)
    kind, kw = _x_format(cfg1, [1, S, S])
    fp = make_format(kind, b0=1, b1=16, **kw)
    vc = v if (v.stride(-1) == 1 and v.stride(0) == S * v.stride(1)) else v.contiguous()
    Vp = torch.empty((B, 3, h, d, S), dtype=torch.bfloat16, device=dev)
    L.check(lib.bq_split3_bf16_transposed(vc.data_ptr(), Vp.data_ptr(), B, S, h, d, vc.stride(1), st), "bq_split3_bf16_transposed")
    scores = torch.empty((B * h, S, S), dtype=torch.float32, device=dev)
    for b in range(B):
        rc = lib.bq_bmm_split_tn(Qq[b].data_ptr(), Kp[b].data_ptr(), scores[b * h].data_ptr(), h, S, S, d, 1, 3, 3, _TA, _TB, S, S * S,
                                 1 if causal else 0, st)        # causal: tiles above the diagonal are never read -> not computed
        L.check(rc, "bq_bmm_split_tn(QK^T)")
    P = torch.empty((B * h, S, S), dtype=torch.bfloat16, device=dev)
    if key_mask is not None and (key_mask.dtype != torch.int32 or key_mask.device != dev or not key_mask.is_contiguous()
                                 or key_mask.shape[0] != B):
        raise ValueError("key_mask must be the int32 [B, words] bitmap of key_mask_bits() on the operands' device")
    # with a key-padding bitmap a row can be fully masked, and the reference then spreads it uniformly over ALL keys, future ones
    # included: only the mask-free causal case may drop the probabilities / products behind the diagonal
    tri = 2 if (causal and key_mask is None) else 0
    rc = lib.bq_softmax_quantize(ctypes.byref(fp), scores.data_ptr(), P.data_ptr(), B * h, h, S, S, S, S * S, S, S * S, float(score_div),
                                 tri if tri else (1 if causal else 0), key_mask.data_ptr() if key_mask is not None else None,
                                 key_mask.shape[1] if key_mask is not None else 0, st)
    L.check(rc, "bq_softmax_quantize")
    del scores
    out = torch.empty((B, S, H), dtype=torch.float32, device=dev)
    for b in range(B):
        rc = lib.bq_bmm_split_tn(P[b * h].data_ptr(), Vp[b].data_ptr(), out[b].data_ptr(), h, S, d, S, 1, 3, 3, _TA, _TB, H, d,
                                 tri, st)                       # causal: the K loop stops at the last visible key of the row tile
        L.check(rc, "bq_bmm_split_tn(PV)")
    return out
