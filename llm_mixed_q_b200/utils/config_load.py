"""TOML quant-config I/O with the reference's "NA" <-> None convention (reference utils/config_load.py:6-59)."""
from pathlib import Path

import toml


def _map_leaves(obj, fn):
    if isinstance(obj, dict):
        for key in obj:
            obj[key] = _map_leaves(obj[key], fn)
        return obj
    if isinstance(obj, list):
        return [_map_leaves(v, fn) for v in obj]
    if isinstance(obj, tuple):
        return tuple(_map_leaves(v, fn) for v in obj)
    return fn(obj)


def convert_str_na_to_none(d):
    """TOML has no null: the string "NA" stands for None."""
    return _map_leaves(d, lambda v: None if isinstance(v, str) and v == "NA" else v)


def convert_none_to_str_na(d):
    return _map_leaves(d, lambda v: "NA" if v is None else v)


def load_config(config_path):
    with open(config_path, "r") as f:
        return convert_str_na_to_none(toml.load(f))


def save_config(config, config_path):
    config = convert_none_to_str_na(config)
    Path(config_path).parent.mkdir(parents=True, exist_ok=True)
    with open(config_path, "w") as f:
        toml.dump(config, f)
