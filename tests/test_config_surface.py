"""CPU: the config surface (TOML keys, per-op filtering, per-model expansion, registries) reproduces the
reference's parser output on every shipped TOML (tests/golden/configs.json)."""
import copy
import json
import os

import pytest

from llm_mixed_q_b200.models.quantize import (QUANTIZED_FUNC_MAP, QUANTIZED_MODULE_MAP, QUANTIZER_MAP, get_quantized_cls,
                                               get_quantized_func, get_quantizer, parse_node_config)
from llm_mixed_q_b200.models.quantize.quant_config_parser import MASE_OP_TO_ENTRIES, QUANT_ARITH_ENTRIES
from llm_mixed_q_b200.utils.config_load import (convert_none_to_str_na, convert_str_na_to_none, load_config, save_config)


def clone(d):
    return json.loads(json.dumps(d))


def test_parse_node_config_matches_reference(golden_configs):
    n = 0
    for node in golden_configs["node"]:
        raw = convert_str_na_to_none(clone(golden_configs["raw"][node["file"]]["default"]))
        if "error" in node:
            with pytest.raises(Exception) as ei:
                parse_node_config(raw, node["op"])
            assert type(ei.value).__name__ == node["error"]
        else:
            assert parse_node_config(raw, node["op"]) == node["parsed"], (node["file"], node["op"])
        n += 1
    assert n >= 40


def test_unknown_op_asserts():
    with pytest.raises(AssertionError):
        parse_node_config({"name": "block_fp"}, "not_an_op")


def test_missing_key_raises_keyerror_and_non_strict_skips():
    cfg = {"name": "block_fp", "is_ptq": True, "data_in_width": 6}
    with pytest.raises(KeyError):
        parse_node_config(cfg, "linear")
    assert parse_node_config(cfg, "linear", strict=False) == cfg


def test_bypass_returned_unfiltered():
    cfg = {"name": "integer", "bypass": True, "junk": 1}
    assert parse_node_config(cfg, "linear") is cfg


def test_opt_expansion_matches_reference(golden_configs):
    from llm_mixed_q_b200.models.opt_quantized import parse_opt_quantized_config

    for fn, raw in golden_configs["raw"].items():
        exp = golden_configs["opt"][fn]
        if "error" in exp:
            with pytest.raises(Exception) as ei:
                parse_opt_quantized_config(clone(raw), 2)
            assert type(ei.value).__name__ == exp["error"]
        else:
            assert parse_opt_quantized_config(clone(raw), 2) == exp, fn
    # section 4.4 style per-layer mixed precision file, unspecified nodes fall back to [default]
    got = parse_opt_quantized_config(clone(golden_configs["mixed_raw"]), 3)
    assert got == golden_configs["mixed_opt"]
    assert got["model_layer_2"]["fc1"]["data_in_width"] == 6       # layer 2 not listed -> default
    assert parse_opt_quantized_config(None, 2) is None


def test_llama_expansion_matches_reference(golden_configs):
    from llm_mixed_q_b200.models.llama_quantized import parse_llama_quantized_config

    for fn, raw in golden_configs["raw"].items():
        exp = golden_configs["llama"][fn]
        if "error" in exp:
            with pytest.raises(Exception) as ei:
                parse_llama_quantized_config(clone(raw), 2)
            assert type(ei.value).__name__ == exp["error"]
        else:
            assert parse_llama_quantized_config(clone(raw), 2) == exp, fn


def test_toml_roundtrip_na(tmp_path, golden_configs):
    cfg = convert_str_na_to_none(clone(golden_configs["mixed_raw"]))
    assert cfg["model_layer_0"]["fc1"]["data_in_exponent_bias"] is None
    p = os.path.join(tmp_path, "sub", "c.toml")
    save_config(copy.deepcopy(cfg), p)
    assert '"NA"' in open(p).read()
    assert load_config(p) == cfg
    assert convert_none_to_str_na({"a": [None, 1], "b": (None,)}) == {"a": ["NA", 1], "b": ("NA",)}


def test_registries_have_reference_names():
    for name in ("block_fp", "block_minifloat", "block_log", "minifloat_denorm", "minifloat_ieee", "integer"):
        assert name in QUANTIZER_MAP
        assert name in QUANTIZED_MODULE_MAP["linear"]
        assert name in QUANTIZED_FUNC_MAP["matmul"] and name in QUANTIZED_FUNC_MAP["bmm"]
        assert name in QUANTIZED_FUNC_MAP["rotary_positional_encoding"]
    # "log" aliases the block_log functions (reference quantized_functions/__init__.py:20,29)
    assert QUANTIZED_FUNC_MAP["bmm"]["log"] is QUANTIZED_FUNC_MAP["bmm"]["block_log"]
    assert QUANTIZED_FUNC_MAP["matmul"]["log"] is QUANTIZED_FUNC_MAP["matmul"]["block_log"]
    cfg = {"name": "block_fp"}
    assert get_quantized_cls("linear", cfg).__name__ == "LinearBlockFP"
    assert get_quantized_func("bmm", cfg).__name__ == "bmm_block_fp"
    assert get_quantizer("linear", cfg).__name__ == "block_fp_quantizer"
    assert set(MASE_OP_TO_ENTRIES) >= {"linear", "matmul", "bmm", "rotary_positional_encoding"}
    assert QUANT_ARITH_ENTRIES["block_fp"]["data_in_entries"] == (
        "data_in_width", "data_in_exponent_width", "data_in_exponent_bias", "data_in_block_size")


def test_quantizer_signatures_match_reference():
    import inspect

    from llm_mixed_q_b200.models.quantize.quantizers import (block_fp_quantizer, block_log_quantizer,
                                                             block_minifloat_quantizer, minifloat_denorm_quantizer)

    def sig(fn):
        return [(p.name, p.default) for p in inspect.signature(fn).parameters.values()]

    E = inspect.Parameter.empty
    assert sig(block_fp_quantizer) == [("x", E), ("width", 12), ("exponent_width", 8), ("exponent_bias", None),
                                       ("block_size", [16]), ("skip_first_dim", True)]
    assert sig(block_minifloat_quantizer) == [("x", E), ("width", E), ("exponent_width", E), ("exponent_bias_width", E),
                                              ("block_size", [16]), ("skip_first_dim", False)]
    assert sig(block_log_quantizer) == [("x", E), ("width", E), ("exponent_bias_width", None), ("block_size", [16]),
                                        ("skip_first_dim", False)]
    assert sig(minifloat_denorm_quantizer) == [("x", E), ("width", E), ("exponent_width", E), ("exponent_bias", None)]


def test_dict_tools_docstring_examples_and_search_roundtrip(golden_configs):
    """flatten/expand known answers are the worked examples in reference utils/dict_tools.py:1-75; the round trip is the
    search's save path (search/search.py:873-897): nested -> "root:..." flat params -> nested -> TOML -> parser."""
    from llm_mixed_q_b200.models.opt_quantized import parse_opt_quantized_config
    from llm_mixed_q_b200.utils.dict_tools import expand_dict, flatten_dict, parse_ast_literal, resolve_ast_literals

    nested = {"a": 1, "b": {"c": 2, "d": {"e": 3, "f": 4}}}
    flat = {}
    flatten_dict(nested, flat, join=":", name="root")
    assert flat == {"root:a": 1, "root:b:c": 2, "root:b:d:e": 3, "root:b:d:f": 4}
    back = {}
    expand_dict(flat, back, join=":", name="root")
    assert back == nested
    with pytest.raises(ValueError):
        expand_dict({"root:a": 1, "root:a:b": 2}, {"a": 1})
    assert parse_ast_literal("!ast![1, 16]") == [1, 16] and parse_ast_literal("!ast!None") is None
    assert parse_ast_literal("block_fp") == "block_fp" and parse_ast_literal(6) == 6
    assert resolve_ast_literals({"x": ["!ast!True", 5], "y": {"z": "!ast!(1, 2)"}}) == {"x": [True, 5], "y": {"z": (1, 2)}}
    mixed = convert_str_na_to_none(clone(golden_configs["mixed_raw"]))
    flat, back = {}, {}
    flatten_dict(mixed, flat)
    assert "root:model_layer_0:self_attn:q_proj:data_in_width" in flat
    expand_dict(flat, back)
    assert parse_opt_quantized_config(back, 3) == golden_configs["mixed_opt"]


def test_shipped_configs_parse():
    from llm_mixed_q_b200.models.llama_quantized import parse_llama_quantized_config
    from llm_mixed_q_b200.models.opt_quantized import parse_opt_quantized_config

    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "configs")
    q = parse_opt_quantized_config(os.path.join(root, "opt_6.7b_mixed_bfp.toml"), 32)
    widths = {q[f"model_layer_{i}"]["fc1"]["weight_width"] for i in range(32)}
    assert widths <= {5, 4, 3, 2} and len(widths) > 1
    assert q["model_layer_0"]["self_attn"]["bmm_0"]["data_in_exponent_bias"] is None
    for name, kind in (("llama_w4a4_block_minifloat.toml", "block_minifloat"), ("llama_w4a4_block_log.toml", "block_log")):
        q = parse_llama_quantized_config(os.path.join(root, name), 2)
        assert q["model_layer_1"]["mlp"]["down_proj"]["name"] == kind
        assert q["model_layer_1"]["mlp"]["down_proj"]["weight_width"] == 4
        assert q["model_layer_0"]["self_attn"]["rotary_positional_encoding"]["name"] == "integer"
