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


def _tiny_opt_layer(block_size):
    from llm_mixed_q_b200.models.opt_quantized import OPTQuantizedConfig
    from llm_mixed_q_b200.models.opt_quantized.modeling_opt import OPTQuantizedDecoderLayer

    d = {"bypass": False, "name": "block_fp", "is_ptq": True}
    for p in ("data_in", "weight", "bias"):
        d.update({f"{p}_width": 6, f"{p}_exponent_width": 8, f"{p}_exponent_bias": None,
                  f"{p}_block_size": [16] if p == "bias" else block_size})
    cfg = OPTQuantizedConfig(hidden_size=256, num_hidden_layers=1, ffn_dim=512, num_attention_heads=4, vocab_size=128,
                             max_position_embeddings=64, quant_config={"default": d})
    return OPTQuantizedDecoderLayer(cfg, 0)


def test_fused_plan_resolves_blocks_against_the_real_operand_shape():
    """reference quantizers/utils.py:42-67: block_size [16] on the 3-D [B, S, H] input of q/k/v/out_proj means [1, S, 16] — a
    block spans every token — so the 1x16 epilogue quantizers must NOT be selected; [1, 16] is the row-block case they serve."""
    from llm_mixed_q_b200.models.quantize.quantized_functions.attention import output_quantizable
    from llm_mixed_q_b200.models.quantize.quantized_functions.fused_glue import linear_input_format, row_block16_format

    lyr = _tiny_opt_layer([1, 16])
    assert lyr._fused_plan(32) is not None
    lyr16 = _tiny_opt_layer([16])
    assert lyr16._fused_plan(32) is None
    at = lyr16.self_attn
    assert linear_input_format(at.q_proj, rows=32) is None              # 3-D input: [1, 32, 16] blocks
    assert linear_input_format(lyr16.fc1) is not None                   # 2-D input (reference :412): [1, 16] blocks
    assert not output_quantizable(at.out_proj.config, 256, 32) and output_quantizable(at.out_proj.config, 256, 1)
    assert row_block16_format(at.quant_config["bmm_0"], "data_in", 64, rows=32) is None
    assert row_block16_format(at.quant_config["bmm_0"], "data_in", 64, rows=1) is not None


def test_ptq_linear_keeps_autograd_semantics_of_the_reference():
    """reference linear.py:63-71 runs F.linear outside no_grad: weight / bias get gradients in PTQ mode; the forward-only fused
    kernels are therefore only eligible when no graph is being recorded (checked before entering no_grad)."""
    import inspect

    from llm_mixed_q_b200.models.quantize.quantized_modules import linear as lin_mod

    src = inspect.getsource(lin_mod._LinearBase.forward)
    assert "wants_grad" in src and src.index("wants_grad =") < src.index("with torch.no_grad()")


def test_llama_gemm_epilogue_eligibility_predicates():
    """Host-side decisions of the Llama GEMM-epilogue fusions (quantized_modules/linear.py): the gated-SiLU epilogue needs two bias-free
    PTQ projections of one shape with whole blocks of 16 features; the RoPE epilogue head sizes 64 / 128; q | k | v in one launch
    additionally 256-column segments and either all or none of the three biases."""
    from llm_mixed_q_b200.models.quantize import get_quantized_cls
    from llm_mixed_q_b200.models.quantize.quantized_modules import linear as QL

    cfg = {"name": "block_fp", "bypass": False, "is_ptq": True}
    for p in ("data_in", "weight", "bias"):
        cfg.update({f"{p}_width": 6, f"{p}_exponent_width": 8, f"{p}_exponent_bias": None, f"{p}_block_size": [1, 16]})
    mk = lambda k, n, bias=False, c=cfg: get_quantized_cls("linear", c)(k, n, bias=bias, config=dict(c))
    assert QL.gated_silu_fusable(mk(64, 352), mk(64, 352))
    assert not QL.gated_silu_fusable(mk(64, 352), mk(64, 368))                  # different shapes
    assert not QL.gated_silu_fusable(mk(64, 352, bias=True), mk(64, 352))       # a bias on one of them
    assert not QL.gated_silu_fusable(mk(64, 40), mk(64, 40))                    # below one 64-column pair of chunks
    qat = dict(cfg, is_ptq=False)
    assert not QL.gated_silu_fusable(mk(64, 352, c=qat), mk(64, 352))           # QAT: weights are re-quantised every call
    wide = dict(cfg, weight_width=12)
    assert not QL.gated_silu_fusable(mk(64, 352, c=wide), mk(64, 352))          # > 8 significant bits: not bf16-exact
    assert QL.rope_epilogue_fusable(mk(64, 256), 128) and QL.rope_epilogue_fusable(mk(64, 256), 64)
    assert not QL.rope_epilogue_fusable(mk(64, 256), 32) and not QL.rope_epilogue_fusable(mk(64, 192), 128)
    q, k, v = mk(64, 256), mk(64, 256), mk(64, 256)
    assert QL.qkv_rope_fusable(q, k, v, 128) and QL.qkv_plain_fusable(q, k, v)
    assert not QL.qkv_rope_fusable(mk(64, 384), mk(64, 384), mk(64, 384), 128)  # segments of 384 columns: a 256-wide tile would straddle two
    assert QL.qkv_plain_fusable(mk(64, 384), mk(64, 384), mk(64, 384))
    assert not QL.qkv_rope_fusable(q, k, mk(64, 256, bias=True), 128) and not QL.qkv_plain_fusable(q, k, mk(64, 256, bias=True))
    try:
        QL.QKV_ONE_LAUNCH = False
        assert not QL.qkv_rope_fusable(q, k, v, 128) and not QL.qkv_plain_fusable(q, k, v)
    finally:
        QL.QKV_ONE_LAUNCH = True
