"""
Llama rotary embedding with quantised cos/sin tables — reference
quantized_functions/rotary_positional_encoding.py:40-248 (adjacent row f3 of SURVEY.md §8).
Only the tables are quantised (q, k untouched); every shipped TOML selects the integer variant.
"""
from functools import partial

import torch

from ..quantized_modules.linear import operand_format
from ..quantizers import QUANTIZER_MAP


def _rotate_half(x):
    half = x.shape[-1] // 2
    return torch.cat((-x[..., half:], x[..., :half]), dim=-1)


def _apply(q, k, cos, sin, position_ids, table_quantizer):
    cos = table_quantizer(cos.squeeze(1).squeeze(0))   # [seq_len, dim]
    sin = table_quantizer(sin.squeeze(1).squeeze(0))
    cos = cos[position_ids].unsqueeze(1)               # [bs, 1, seq_len, dim]
    sin = sin[position_ids].unsqueeze(1)
    return (q * cos) + (_rotate_half(q) * sin), (k * cos) + (_rotate_half(k) * sin)


def apply_token_major(q, k, cos, sin, position_ids, config):
    """Same arithmetic as `_apply` on token-major operands: q, k [B, S, h, d] (views of the [B, S, H] projections) instead of
    [B, h, S, d].  Element-wise, so the results are bit-identical to the head-major call; it saves the two transposes the
    fused attention kernel does not need (it reads [B, S, H] directly)."""
    tq = _table_quantizer(config, config["name"])
    cos = tq(cos.squeeze(1).squeeze(0))
    sin = tq(sin.squeeze(1).squeeze(0))
    cos = cos[position_ids].unsqueeze(2)               # [bs, seq_len, 1, dim]
    sin = sin[position_ids].unsqueeze(2)
    return (q * cos) + (_rotate_half(q) * sin), (k * cos) + (_rotate_half(k) * sin)


def apply_token_major_quantized(q, k, cos, sin, position_ids, config, matmul0_config, num_heads):
    """RoPE on token-major q / k [B, S, H] fp32 FUSED with matmul_0's operand quantizers (bq_rope_quantize): returns the bf16
    operands (Qq, Kq) of the attention kernel, or None when the formats are not the [1,16] block_fp / block_minifloat case the kernel
    serves (the caller then runs apply_token_major + quantize_qkv — same results, ~14 launches instead of 2)."""
    import ctypes

    from .... import _lib as L

    B, S, H = q.shape
    d = H // num_heads
    if not (q.is_cuda and q.dtype == torch.float32 and k.dtype == torch.float32):
        return None
    prep = rope_quantize_operands(cos, sin, position_ids, config, matmul0_config, B, S, d)
    if prep is None:
        return None
    cos_t, sin_t, pos, fq, fk = prep
    qc, kc = q, k
    if qc.stride(-1) != 1 or qc.stride(0) != S * qc.stride(1):
        qc = qc.contiguous()
    if kc.stride(-1) != 1 or kc.stride(0) != S * kc.stride(1):
        kc = kc.contiguous()
    Qq = torch.empty((B, S, H), dtype=torch.bfloat16, device=q.device)
    Kq = torch.empty((B, S, H), dtype=torch.bfloat16, device=q.device)
    rc = L.load().bq_rope_quantize(qc.data_ptr(), kc.data_ptr(), cos_t.data_ptr(), sin_t.data_ptr(),
                                   pos.data_ptr() if pos is not None else None, cos_t.shape[0], B, S, num_heads, d,
                                   qc.stride(1), kc.stride(1), ctypes.byref(fq), ctypes.byref(fk), Qq.data_ptr(), Kq.data_ptr(),
                                   L.stream_ptr(q.device))
    L.check(rc, "bq_rope_quantize")
    return Qq, Kq


def rope_quantize_operands(cos, sin, position_ids, config, matmul0_config, B, S, d):
    """What the fused RoPE kernels need besides q / k: the quantised [table_rows, d] cos / sin tables (quantised exactly as the
    reference quantises them, :27-36), the validated int64 [B, S] positions (None for the default arange) and the formats of
    matmul_0's x (q, blocks along d) and y (k^T, blocks along S) operands.  None when the configuration is not the [1,16]
    block_fp / block_minifloat case those kernels serve."""
    from ..quantizers.utils import make_format, resolve_block_shape

    if d % 32 != 0 or S % 16 != 0:
        return None
    (qk_, qkw, qbs), (kk_, kkw, kbs) = operand_format(matmul0_config, "data_in"), operand_format(matmul0_config, "weight")
    if qk_ not in ("block_fp", "block_minifloat") or kk_ not in ("block_fp", "block_minifloat") or qbs is None or kbs is None:
        return None
    if resolve_block_shape([1, S, d], qbs)[1:] != [1, 16] or resolve_block_shape([1, d, S], kbs)[1:] != [1, 16]:
        return None
    tq = _table_quantizer(config, config["name"])
    cos_t = tq(cos.squeeze(1).squeeze(0)).contiguous()          # [seq_len, d], quantised exactly as the reference does
    sin_t = tq(sin.squeeze(1).squeeze(0)).contiguous()
    if cos_t.dtype != torch.float32 or cos_t.shape[-1] != d or not cos_t.is_cuda:
        return None
    pos = None
    if position_ids is not None:
        # explicit positions index the [table_rows, d] tables like the reference's cos[position_ids] (:44-45): out-of-range
        # raises IndexError, negative indices wrap.  One device-to-host look — the model classes pass None for the default
        # arange, so the captured / steady-state forward never takes it.  (The kernels additionally clamp: no OOB read.)
        rows = cos_t.shape[0]
        pos = position_ids.expand(B, S).to(torch.int64)
        lo, hi = (int(v) for v in torch.stack((pos.min(), pos.max())).tolist())
        if lo < -rows or hi >= rows:
            raise IndexError(f"index {hi if hi >= rows else lo} is out of bounds for dimension 0 with size {rows}")
        if lo < 0:
            pos = torch.where(pos < 0, pos + rows, pos)
        pos = pos.contiguous()
    return cos_t, sin_t, pos, make_format(qk_, b0=1, b1=16, **qkw), make_format(kk_, b0=1, b1=16, **kkw)


def _table_quantizer(config, name):
    if config.get("bypass", False):
        return lambda t: t
    if name == "integer":
        return partial(QUANTIZER_MAP["integer"], width=config["data_in_width"], frac_width=config["data_in_frac_width"])
    kind, kw, block_size = operand_format({**config, "name": name}, "data_in")
    fn = QUANTIZER_MAP[kind]
    if block_size is not None:
        return partial(fn, block_size=block_size, skip_first_dim=False, **kw)
    return partial(fn, **kw)


def _make(name):
    def apply_rotary_pos_emb(q, k, cos, sin, position_ids, config):
        return _apply(q, k, cos, sin, position_ids, _table_quantizer(config, name))

    apply_rotary_pos_emb.__name__ = f"apply_rotary_pos_emb_{name}"
    return apply_rotary_pos_emb


apply_rotary_pos_emb_integer = _make("integer")
apply_rotary_pos_emb_block_fp = _make("block_fp")
apply_rotary_pos_emb_block_minifloat = _make("block_minifloat")
apply_rotary_pos_emb_block_log = _make("block_log")
apply_rotary_pos_emb_minifloat_denorm = _make("minifloat_denorm")
apply_rotary_pos_emb_minifloat_ieee = _make("minifloat_ieee")
