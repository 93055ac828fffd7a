"""
Per-model quant-config expansion, shared by OPT / Llama / BERT.

The reference writes the same expansion three times (opt_quantized/quant_config_opt.py:34-97,
llama_quantized/quant_config_llama.py:38-116, bert_quantized/quant_config_bert.py).  Here a model is
described by a *layer template* — a nested dict whose leaves name the op kind ("linear" / "matmul" /
"rotary_positional_encoding") — and one generic routine produces the identical nested result:

    {"model_layer_<i>": {<template with every leaf replaced by its parsed node config>}, "default": {...}}

Lookup order for a leaf (same as the reference): `[model_layer_<i>.<path>]` → `[model_layer.<path>]` →
the type-level section (`[linear]`, `[bmm]`/`[matmul]`, `[rotary_positional_encoding]`) → `[default]`.
"""
from __future__ import annotations

from copy import deepcopy

import toml

from ..utils.config_load import convert_str_na_to_none
from .quantize.quant_config_parser import parse_node_config


def _expand(template: dict, layer_qc: dict, type_defaults: dict, strict: bool) -> dict:
    out = {}
    for key, leaf in template.items():
        if isinstance(leaf, dict):
            out[key] = _expand(leaf, layer_qc.get(key, {}) if isinstance(layer_qc, dict) else {}, type_defaults, strict)
        else:
            op = leaf
            node = layer_qc.get(key, type_defaults[op]) if isinstance(layer_qc, dict) else type_defaults[op]
            out[key] = deepcopy(parse_node_config(node, op, strict=strict))
    return out


def parse_model_quant_config(config, num_hidden_layers: int, template: dict, type_sections: dict, strict: bool = True):
    """
    config: TOML path | dict | None.  type_sections: op kind -> TOML section holding its default
    (e.g. {"linear": "linear", "matmul": "bmm"} for OPT).
    """
    assert isinstance(config, (str, dict, type(None))), "Must provide either a path, None or a dict"
    if config is None:
        return None
    if isinstance(config, str):
        config = toml.load(config)
    config = convert_str_na_to_none(config)
    assert "default" in config, "Must provide default config"
    default_qc = config["default"]
    type_defaults = {op: parse_node_config(config.get(section, default_qc), mase_op=op) for op, section in type_sections.items()}
    general_layer_qc = config.get("model_layer", None)
    parsed = {}
    for i in range(num_hidden_layers):
        entry = f"model_layer_{i}"
        layer_qc = config.get(entry, general_layer_qc)
        parsed[entry] = _expand(template, layer_qc if layer_qc is not None else {}, type_defaults, strict)
    parsed["default"] = default_qc
    return parsed
