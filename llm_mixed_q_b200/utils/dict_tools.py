"""Flat <-> nested quant-config dicts and the search files' "!ast!" literals.

The reference's mixed-precision search records a sampled per-layer config as flat optuna parameters named
"root:model_layer_3:self_attn:q_proj:data_in_width" and re-nests them before `save_config` writes the TOML
(reference utils/dict_tools.py:1-89, search/search.py:873-897); list-valued choices in the search-space TOMLs are strings
such as "!ast![1, 16]" evaluated with `ast.literal_eval` (quant_config_sampler.py:13-14).  Same argument order and
in-place `new_d` convention as the reference, so configs written by its search load unchanged.
"""
from __future__ import annotations

import ast

AST_PREFIX = "!ast!"


def flatten_dict(d: dict, new_d: dict, join: str = ":", name: str = "root") -> None:
    for key, value in d.items():
        path = f"{name}{join}{key}"
        if isinstance(value, dict):
            flatten_dict(value, new_d, join, path)
        else:
            new_d[path] = value


def expand_dict(d: dict, new_d: dict, join: str = ":", name: str = "root") -> None:
    prefix = f"{name}{join}"
    for flat_key, value in d.items():
        keys = flat_key.removeprefix(prefix).split(join)
        node = new_d
        for k in keys[:-1]:
            node = node.setdefault(k, {})
        leaf = keys[-1]
        if leaf not in node:
            node[leaf] = value
        elif isinstance(node[leaf], dict):
            node[leaf].update(value)
        else:
            raise ValueError(f"Cannot create nested dict at {keys} with value {value}")


def parse_ast_literal(value):
    """"!ast![1, 16]" -> [1, 16]; "!ast!None" -> None; anything else is returned unchanged."""
    if isinstance(value, str) and value.startswith(AST_PREFIX):
        return ast.literal_eval(value.removeprefix(AST_PREFIX))
    return value


def resolve_ast_literals(obj):
    """Applies `parse_ast_literal` to every leaf of a nested dict / list (returns a new structure)."""
    if isinstance(obj, dict):
        return {k: resolve_ast_literals(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(resolve_ast_literals(v) for v in obj)
    return parse_ast_literal(obj)
