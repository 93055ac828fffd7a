"""CPU: the oracle restatement (oracle/oracle.py) is pinned bit-exactly to the reference's own outputs
(tests/golden/, produced by oracle/gen_golden.py running the unmodified reference)."""
import hashlib

import numpy as np
import pytest
import torch

from conftest import bits_equal, case_input, case_kwargs, f32
from golden_inputs import hashed_input
from oracle import oracle as O

BLOCKED = ("block_fp", "block_minifloat", "block_log")


def run_oracle(case, x):
    fn = O.QUANTIZERS[case["fmt"]]
    kw = case_kwargs(case)
    if case["fmt"] in BLOCKED:
        return fn(x, block_size=list(case["block_size"]), skip_first_dim=case["skip_first_dim"], **kw)
    return fn(x, **kw)


def test_oracle_matches_reference_goldens(golden_quantizers):
    arrays, cases = golden_quantizers
    assert len(cases) > 1500
    bad = []
    for case in cases:
        y = run_oracle(case, case_input(arrays, case))
        if not bits_equal(y, f32(arrays[case["key"]]).reshape(y.shape)):
            bad.append((case["key"], case["fmt"], case["layout"], case["input"], case["block_size"]))
    assert not bad, f"{len(bad)} oracle/reference mismatches, first: {bad[:5]}"


def test_oracle_hashed_large_cases(golden_hashed):
    for case in golden_hashed:
        x = hashed_input(case)
        fn = O.QUANTIZERS[case["fmt"]]
        if case["block_size"] is not None:
            y = fn(x, block_size=case["block_size"], skip_first_dim=case["skip_first_dim"], **case["kwargs"])
        else:
            y = fn(x, **case["kwargs"])
        h = hashlib.sha256(y.contiguous().view(torch.int32).numpy().tobytes()).hexdigest()
        assert h == case["sha256"], (case["tag"], case["fmt"], case["kwargs"])


def test_oracle_docstring_known_answers():
    # reference minifloat.py:41-43: 1 0111 011 (4 exponent / 3 mantissa bits, bias 15) = -0.00146484375
    y = O.minifloat_denorm_quantize(torch.tensor([-0.00146484375]), 8, 4, 15)
    assert float(y) == -0.00146484375
    # reference minifloat.py:150-153: same bits with the implicit one = -0.00537109375
    y = O.minifloat_ieee_quantize(torch.tensor([-0.00537109375]), 8, 4, 15)
    assert float(y) == -0.00537109375


def test_oracle_consumers_match_reference(golden_consumers):
    arrays, cases = golden_consumers
    for c in cases:
        k = c["key"]
        if c["op"] == "linear":
            x, w, b = f32(arrays[k + "_x"]), f32(arrays[k + "_w"]), f32(arrays[k + "_b"])
            y, wq, bq = O.linear_forward(x, w, b, c["config"])
            assert bits_equal(wq, f32(arrays[k + "_wq"])), k          # in-place PTQ overwrite values
            assert bits_equal(bq, f32(arrays[k + "_bq"])), k
            torch.testing.assert_close(y, f32(arrays[k + "_y"]), rtol=1e-5, atol=1e-5)
        else:
            x, yb = f32(arrays[k + "_x"]), f32(arrays[k + "_y"])
            yv = yb.transpose(-1, -2) if c["y_transposed"] else yb
            out = O.matmul_forward(x, yv, c["config"], style=c["op"])
            torch.testing.assert_close(out, f32(arrays[k + "_o"]), rtol=1e-5, atol=1e-4)


def test_perplexity_reduction():
    # eval/eval_lm.py:41-63: exp(sum(loss*B*S)/(S*N))
    ppl = O.perplexity_from_losses([2.0, 4.0], batch_size=3, seq_len=7)
    assert abs(ppl - np.exp(3.0)) < 1e-9
