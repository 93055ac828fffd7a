"""CPU: the analytic layer / model profiler reproduces the unmodified reference's numbers and failure modes on
every shipped quantization TOML and on the mixed-precision config (tests/golden/profiler.json, oracle/gen_golden_profiler.py)."""
import json
import os
from copy import deepcopy

import numpy as np
import pytest

from conftest import GOLD


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(GOLD, "profiler.json")) as f:
        p = json.load(f)
    with open(os.path.join(GOLD, "configs.json")) as f:
        c = json.load(f)
    return p, c


def _run(fn, want):
    if "error" in want:
        with pytest.raises(Exception) as ei:
            fn()
        assert type(ei.value).__name__ == want["error"]
    else:
        got = fn()
        assert {k: int(v) for k, v in got.items()} == want
        assert all(isinstance(v, (np.integer, int)) for v in got.values())


def test_linear_and_matmul_layers_match_reference(gold):
    from llm_mixed_q_b200.models.quantize import profile_linear_layer, profile_matmul_layer

    p, c = gold
    assert len(p["linear"]) >= 30 and len(p["matmul"]) >= 30
    for case in p["linear"]:
        node = deepcopy(c["raw"][case["toml"]]["default"])
        fin, fout, bias, bs = case["args"]
        _run(lambda: profile_linear_layer(node, fin, fout, bias, bs), case["result"])
    for case in p["matmul"]:
        node = deepcopy(c["raw"][case["toml"]]["default"])
        s0, s1 = (tuple(a) for a in case["args"])
        _run(lambda: profile_matmul_layer(node, s0, s1), case["result"])


def test_model_profilers_match_reference(gold):
    from llm_mixed_q_b200.models import get_config_cls, get_model_profiler

    p, c = gold
    seen = set()
    for case in p["models"]:
        raw = deepcopy(c["raw"][case["toml"]]) if case["toml"] else deepcopy(c["mixed_raw"])
        seen.add(case["arch"])

        def run():
            cfg = get_config_cls(case["arch"])(quant_config=raw, **case["kw"])
            return get_model_profiler(case["arch"])(cfg, case["seq_len"])

        _run(run, case["result"])
    assert seen == {"opt", "llama", "bert"}


def test_update_profile_accumulates_in_place():
    from llm_mixed_q_b200.models.quantize import update_profile

    a = {"num_params": 1, "num_acts": 2, "param_bits": 3, "act_bits": 4, "flops": 5}
    b = update_profile(a, {"num_params": 10, "num_acts": 20, "param_bits": 30, "act_bits": 40, "flops": 50})
    assert b is a and a == {"num_params": 11, "num_acts": 22, "param_bits": 33, "act_bits": 44, "flops": 55}
