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
