"""
Quant-config key schema and per-op filtering — host mirror of reference quant_config_parser.py:32-304.

A config node is a flat dict (`name`, `bypass`, `is_ptq`, `<operand>_<field>` …).  `parse_node_config(cfg, op)`
keeps exactly the keys the op consumes, raising KeyError for a missing required key (strict) and
AssertionError for an unknown op, like the reference.  The schema is expressed as data: per arithmetic, the
field suffixes every operand carries; per op, which operand groups are required / optional.
"""
from __future__ import annotations

from copy import deepcopy

# field suffixes per arithmetic (reference QUANT_ARITH_ENTRIES, quant_config_parser.py:32-155)
_FIELDS = {
    "integer": ("width", "frac_width"),
    "minifloat_ieee": ("width", "exponent_width", "exponent_bias"),
    "minifloat_denorm": ("width", "exponent_width", "exponent_bias"),
    "log": ("width", "exponent_bias"),
    "block_fp": ("width", "exponent_width", "exponent_bias", "block_size"),
    "block_minifloat": ("width", "exponent_width", "exponent_bias_width", "block_size"),
    "block_log": ("width", "exponent_bias_width", "block_size"),
}
_OPERANDS = ("weight", "data_in", "bias", "data_out")

QUANT_ARITH_ENTRIES = {
    arith: {f"{operand}_entries": tuple(f"{operand}_{f}" for f in fields) for operand in _OPERANDS}
    for arith, fields in _FIELDS.items()
}

# <op>: (required groups, optional groups)  (reference MASE_OP_TO_ENTRIES, quant_config_parser.py:236-267)
MASE_OP_TO_ENTRIES = {
    "add": (("name", "data_in_entries"), ("bypass",)),
    "bmm": (("name", "data_in_entries", "weight_entries"), ("bypass",)),
    "conv1d": (("name", "is_ptq", "data_in_entries", "weight_entries"), ("bias_entries", "bypass")),
    "conv2d": (("name", "is_ptq", "data_in_entries", "weight_entries"), ("bias_entries", "bypass")),
    "matmul": (("name", "data_in_entries", "weight_entries"), ("bypass",)),
    "mul": (("name", "data_in_entries"), ("bypass",)),
    "linear": (("name", "is_ptq", "data_in_entries", "weight_entries"), ("bias_entries", "data_out_entries", "bypass")),
    "relu": (("name", "data_in_entries"), ("bypass",)),
    "rotary_positional_encoding": (("name", "data_in_entries"), ("bypass",)),
    "sub": (("name", "data_in_entries"), ("bypass",)),
}


def _copy_keys(src: dict, dst: dict, keys, strict: bool):
    for key in keys:
        if key not in src and not strict:
            continue
        dst[key] = deepcopy(src[key])          # KeyError on a missing required key, like the reference


def _group_keys(arith: str, group: str):
    if group in ("name", "bypass", "is_ptq"):
        return (group,)
    return QUANT_ARITH_ENTRIES[arith][group]


def optional_entry_exists(config: dict, entry_name: str) -> bool:
    prefix = entry_name.removesuffix("_entries")
    return any(key.startswith(prefix) for key in config)


def parse_node_config(config: dict, mase_op: str, strict: bool = True) -> dict:
    """Filter a flat config dict down to what `mase_op` needs (reference quant_config_parser.py:278-304)."""
    assert mase_op in MASE_OP_TO_ENTRIES, f"Unknown mase op: {mase_op}"
    if config.get("bypass", False):
        return config                           # returned unfiltered (reference :287-288)
    required, optional = MASE_OP_TO_ENTRIES[mase_op]
    arith = config["name"]
    parsed = {}
    for group in required:
        _copy_keys(config, parsed, _group_keys(arith, group), strict)
    for group in optional:
        if optional_entry_exists(config, group):
            _copy_keys(config, parsed, _group_keys(arith, group), strict)
    return parsed
