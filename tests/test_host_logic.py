"""CPU: host-side logic of the Python mirror (block-shape inference, layout canonicalisation, exceptions)."""
import pytest
import torch

from llm_mixed_q_b200.models.quantize.quantizers.utils import canonicalise, default_bias, resolve_block_shape
from oracle import oracle as O


@pytest.mark.parametrize("shape", [(48,), (10,), (7, 40), (64, 64), (3, 5, 40), (2, 16, 64)])
@pytest.mark.parametrize("block", [[16], [1, 16], [2, 16], [16, 16], [4], [3, 5], [1, 1, 16], [-1, 16], [128]])
def test_block_shape_inference_matches_oracle(shape, block):
    assert resolve_block_shape(shape, block) == O.infer_block_shape(list(shape), list(block))


def test_canonicalise_layouts():
    x = torch.zeros(6, 40)
    c = canonicalise(x, [1, 16], True)
    assert (c.L, c.R, c.C, c.b0, c.b1, c.fold) == (6, 1, 40, 1, 16, False)
    c = canonicalise(x, [2, 16], False)
    assert (c.L, c.R, c.C, c.b0, c.b1, c.fold) == (1, 6, 40, 2, 16, True)
    c = canonicalise(torch.zeros(48), [1, 16], False)
    assert (c.L, c.R, c.C, c.b0, c.b1, c.fold) == (1, 1, 48, 1, 16, False)
    x3 = torch.zeros(3, 8, 48).transpose(1, 2)           # k^T view
    c = canonicalise(x3, [1, 16], True)
    assert (c.L, c.R, c.C, c.sL, c.sR, c.sC, c.b0, c.b1, c.fold) == (3, 48, 8, 384, 1, 48, 1, 8, True)
    c = canonicalise(torch.zeros(3, 5, 40), [16], True)  # [16] on 3-D: whole second dim in a block
    assert (c.b0, c.b1) == (5, 16)


def test_reference_exceptions_are_mirrored():
    with pytest.raises(RuntimeError):
        canonicalise(torch.zeros(2, 2, 2, 2), [16], True)
    with pytest.raises(NotImplementedError):
        canonicalise(torch.zeros(2, 2, 2), [16], False)
    with pytest.raises(AssertionError):
        canonicalise(torch.zeros(8), [16], True)


def test_default_bias_rule():
    assert default_bias(None, 8) == 127 and default_bias("none", 4) == 7 and default_bias("None", 2) == 1
    assert default_bias(5, 8) == 5


def test_linear_module_surface():
    from llm_mixed_q_b200.models.quantize import get_quantized_cls

    cfg = {"name": "block_fp", "bypass": False, "is_ptq": True}
    for p in ("data_in", "weight", "bias"):
        cfg.update({f"{p}_width": 6, f"{p}_exponent_width": 8, f"{p}_exponent_bias": None, f"{p}_block_size": [1, 16]})
    lin = get_quantized_cls("linear", cfg)(32, 16, bias=True, config=cfg)
    assert isinstance(lin, torch.nn.Linear)
    for attr in ("config", "bypass", "is_ptq", "weight_requires_quantisation", "x_quantizer", "w_quantizer", "b_quantizer"):
        assert hasattr(lin, attr)
    assert lin.weight_requires_quantisation is True and lin.bypass is False
    assert "x/w/b-width=6/6/6" in repr(lin)
    ref = torch.nn.Linear(32, 16)
    lin2 = type(lin).from_float(ref, cfg)
    assert torch.equal(lin2.weight, ref.weight) and torch.equal(lin2.bias, ref.bias)
    byp = get_quantized_cls("linear", cfg)(32, 16, config={"name": "block_fp", "bypass": True})
    x = torch.randn(3, 32)
    assert torch.equal(byp(x), torch.nn.functional.linear(x, byp.weight, byp.bias))   # bypass = plain fp32 linear
