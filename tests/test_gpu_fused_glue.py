"""GPU parity of the glue-fusion kernels: quantising GEMM epilogue, LayerNorm/RMSNorm+quantize, and the
8-kernel fused OPT decoder layer.

The fused kernels apply the reference's ops in the reference's order on the SAME accumulator values the unfused kernels
produce, so epilogue results are compared BIT-EXACTLY against (plain GEMM -> torch ops -> standalone quantizer); the only
stated exception are |x| <= 1e-8 pass-through elements, which the bf16 carrier rounds (see test_gpu_consumers.py).
LayerNorm statistics are summed in a different order than torch's kernel (<= 1-2 ulp in mean / rstd), so a normalised
value can differ by an ulp before quantisation and, on a rounding boundary, by one quantisation step after it."""
import ctypes

import pytest
import torch

from oracle import opt_ref
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def bfp_cfg(width=6):
    d = {"name": "block_fp", "bypass": False, "is_ptq": True}
    for p in ("data_in", "weight", "bias"):
        d.update({f"{p}_width": width, f"{p}_exponent_width": 8, f"{p}_exponent_bias": 127,
                  f"{p}_block_size": [16] if p == "bias" else [1, 16]})
    return d


def assert_bf16_carrier_equal(got_bf16, want_f32, pre_quant_f32):
    passthrough = pre_quant_f32.abs() <= 1e-8
    assert torch.equal(got_bf16.float()[~passthrough], want_f32[~passthrough])
    assert torch.equal(got_bf16[passthrough], want_f32[passthrough].to(torch.bfloat16))


@pytest.mark.parametrize("M,N,K", [(256, 128, 64), (2048, 2048, 512), (1040, 96, 328), (16, 32, 8)])
def test_gemm_epilogue_matches_unfused_composition(M, N, K):
    from llm_mixed_q_b200 import _lib as L
    from llm_mixed_q_b200.models.quantize.quantizers import block_fp_quantizer
    from llm_mixed_q_b200.models.quantize.quantizers.utils import make_format

    lib = L.load()
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    B = (torch.randn(N, K, device="cuda", generator=g) * 0.1).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    res = torch.randn(M, N, device="cuda", generator=g)
    plain = torch.empty(M, N, device="cuda")
    L.check(lib.bq_gemm_bf16_tn(A.data_ptr(), B.data_ptr(), plain.data_ptr(), bias.data_ptr(), 1, M, N, K, K, K, N, 0, 0, 0,
                                L.stream_ptr()), "gemm")

    def run(out_dtype, scale=1.0, act=0, residual=None, fmt=None, qdir=0, use_bias=True):
        ep = L.BqGemmEpilogue()
        ep.bias = bias.data_ptr() if use_bias else None
        if residual is not None:
            ep.residual, ep.ldr = residual.data_ptr(), N
        ep.scale, ep.act, ep.out_dtype = scale, act, (L.BQ_BF16 if out_dtype == torch.bfloat16 else L.BQ_F32)
        if fmt is not None:
            ep.qfmt, ep.qdir = ctypes.pointer(fmt), qdir
        C = torch.full((M, N), float("nan"), device="cuda").to(out_dtype)
        L.check(lib.bq_gemm_bf16_tn_ex(A.data_ptr(), B.data_ptr(), C.data_ptr(), ctypes.byref(ep), M, N, K, K, K, N,
                                       L.stream_ptr()), "gemm_ex")
        return C

    f6 = make_format("block_fp", width=6, exponent_width=8, exponent_bias=127, b0=1, b1=16)
    # 1. bias + residual, fp32 out (out_proj / fc2 of the fused layer)
    assert torch.equal(run(torch.float32, residual=res), res + plain)
    # 2. bias, scale, quantise along N, bf16 (q_proj)
    pre = plain * 0.125
    assert_bf16_carrier_equal(run(torch.bfloat16, scale=0.125, fmt=f6), block_fp_quantizer(pre, 6, 8, 127, [1, 16], True), pre)
    # 3. bias, ReLU, quantise along N, bf16 (fc1 -> fc2)
    pre = torch.relu(plain)
    assert_bf16_carrier_equal(run(torch.bfloat16, act=1, fmt=f6), block_fp_quantizer(pre, 6, 8, 127, [1, 16], True), pre)
    # 4. quantise along M: 16 consecutive rows at one column (k_proj -> k^T operand of bmm_0)
    if M % 16 == 0:
        want = block_fp_quantizer(plain.t().contiguous(), 6, 8, 127, [1, 16], True).t()
        assert_bf16_carrier_equal(run(torch.bfloat16, fmt=f6, qdir=1), want, plain)
    # 5. fp32 output of quantised values (no carrier rounding at all), W4
    f4 = make_format("block_fp", width=4, exponent_width=8, exponent_bias=127, b0=1, b1=16)
    got = run(torch.float32, fmt=f4, use_bias=True)
    want = block_fp_quantizer(plain, 4, 8, 127, [1, 16], True)
    assert torch.equal(got, want)
    # 6. block_minifloat epilogue
    fm = make_format("block_minifloat", width=8, exponent_width=4, exponent_bias_width=8, b0=1, b1=16)
    pre = plain * 8.0
    got = run(torch.float32, scale=8.0, fmt=fm)
    assert torch.equal(got, O.block_minifloat_quantize(pre, 8, 4, 8, [1, 16], True))
    if M % 16 == 0:
        # 7. block_minifloat along M (fp32 out: no carrier rounding), and the rare per-element path: zero blocks, an inf, a
        #    block whose maximum sits within the log2 cliff zone is still on the fast path; NaN/inf blocks take the literal one
        got = run(torch.float32, scale=8.0, fmt=fm, qdir=1)
        want = O.block_minifloat_quantize(pre.t().contiguous(), 8, 4, 8, [1, 16], True).t()
        assert torch.equal(got, want)
        got = run(torch.float32, fmt=f6, qdir=1)
        want = block_fp_quantizer(plain.t().contiguous(), 6, 8, 127, [1, 16], True).t()
        assert torch.equal(got, want)


def test_gemm_epilogue_row_blocks_rare_path_with_nonfinite_values():
    """Quantising along M with blocks that cannot take the fast path (inf / NaN block maxima) and all-zero blocks: the
    per-element out-of-line path must give what the quantizer gives on the transposed fp32 result."""
    from llm_mixed_q_b200 import _lib as L
    from llm_mixed_q_b200.models.quantize.quantizers import block_fp_quantizer
    from llm_mixed_q_b200.models.quantize.quantizers.utils import make_format

    lib = L.load()
    M, N, K = 256, 128, 64
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    B = (torch.randn(N, K, device="cuda", generator=g) * 0.1).to(torch.bfloat16)
    A[3, 5] = float("inf")            # row 3 of C: +-inf
    A[40, 1] = float("nan")           # row 40: NaN
    A[64:80] = 0                      # 16 zero rows: all-zero blocks at every column
    plain = torch.empty(M, N, device="cuda")
    L.check(lib.bq_gemm_bf16_tn(A.data_ptr(), B.data_ptr(), plain.data_ptr(), None, 1, M, N, K, K, K, N, 0, 0, 0, L.stream_ptr()), "gemm")
    f6 = make_format("block_fp", width=6, exponent_width=8, exponent_bias=127, b0=1, b1=16)
    ep = L.BqGemmEpilogue()
    ep.scale, ep.act, ep.out_dtype = 1.0, 0, L.BQ_F32
    ep.qfmt, ep.qdir = ctypes.pointer(f6), 1
    C = torch.full((M, N), 7.0, device="cuda")
    L.check(lib.bq_gemm_bf16_tn_ex(A.data_ptr(), B.data_ptr(), C.data_ptr(), ctypes.byref(ep), M, N, K, K, K, N, L.stream_ptr()), "gemm_ex")
    want = block_fp_quantizer(plain.t().contiguous(), 6, 8, 127, [1, 16], True).t()
    assert torch.equal(torch.isnan(C), torch.isnan(want))
    assert torch.equal(torch.nan_to_num(C, nan=1.0), torch.nan_to_num(want, nan=1.0))
    assert bool((C[64:80] == 0).all())


@pytest.mark.parametrize("warp_rows", [1, 0])
@pytest.mark.parametrize("H,rows", [(2048, 512), (768, 300), (4096, 64), (64, 1000), (1024, 1), (2048, 8 * 148 * 2 + 3),
                                    (4096, 8 * 148 + 5), (3072, 129), (2304, 50), (5120, 40)])
def test_layernorm_quantize_vs_torch_layernorm_then_quantizer(H, rows, warp_rows):
    """Both norm+quantize kernels: row-per-warp / row-per-warp-pair (H <= 2048 / <= 4096, default) and row-per-CTA (any H; forced by
    the A/B switch)."""
    from llm_mixed_q_b200 import _lib as L
    from llm_mixed_q_b200.models.quantize.quantized_functions.fused_glue import norm_quantize
    from llm_mixed_q_b200.models.quantize.quantizers import block_fp_quantizer

    L.load().bq_set_norm_warp_rows(warp_rows)
    try:
        _layernorm_quantize_case(H, rows, norm_quantize, block_fp_quantizer)
    finally:
        L.load().bq_set_norm_warp_rows(1)


def _layernorm_quantize_case(H, rows, norm_quantize, block_fp_quantizer):

    g = torch.Generator(device="cuda").manual_seed(H + rows)
    x = torch.randn(rows, H, device="cuda", generator=g) * 3 + 0.5
    w = 1 + 0.1 * torch.randn(H, device="cuda", generator=g)
    b = 0.1 * torch.randn(H, device="cuda", generator=g)
    f6 = ("block_fp", dict(width=6, exponent_width=8, exponent_bias=127))
    f4 = ("block_fp", dict(width=4, exponent_width=8, exponent_bias=127))
    y6, y4, y6b = norm_quantize(x, w, b, 1e-5, [f6, f4, f6])
    assert y6b.data_ptr() == y6.data_ptr()                      # identical formats share one output
    ln = torch.nn.functional.layer_norm(x, (H,), w, b, 1e-5)
    for got, width in ((y6, 6), (y4, 4)):
        want = block_fp_quantizer(ln, width, 8, 127, [1, 16], True)
        diff = (got.float() - want).abs()
        bad = diff > 0
        frac = float(bad.float().mean())
        assert frac <= 2e-3, frac                                # rounding-boundary flips only
        # a flip is one quantisation step of its block: 2^(E - m) <= 2 * blockmax * 2^-m
        blockmax = ln.abs().view(rows, H // 16, 16).amax(-1, keepdim=True).expand(rows, H // 16, 16).reshape(rows, H)
        assert bool((diff <= 2.0 * blockmax * 2.0 ** -(width - 1) + 1e-8).all())
    # RMSNorm (Llama): y = w * (x * rsqrt(mean(x^2) + eps))
    (r6,) = norm_quantize(x, w, None, 1e-6, [f6])
    var = x.pow(2).mean(-1, keepdim=True)
    rn = w * (x * torch.rsqrt(var + 1e-6))
    want = block_fp_quantizer(rn, 6, 8, 127, [1, 16], True)
    assert float(((r6.float() - want).abs() > 0).float().mean()) <= 2e-3
    # block_minifloat outputs (Llama W4A4 / W8A8), zero rows and a zero block inside a row
    x2 = x.clone()
    x2[0] = 0
    x2[-1, :16] = 0
    fm = ("block_minifloat", dict(width=8, exponent_width=4, exponent_bias_width=8))
    (m8,) = norm_quantize(x2, w, None, 1e-6, [fm])
    rn2 = w * (x2 * torch.rsqrt(x2.pow(2).mean(-1, keepdim=True) + 1e-6))
    want = O.block_minifloat_quantize(rn2, 8, 4, 8, [1, 16], True)
    assert float(((m8.float() - want).abs() > 0).float().mean()) <= 2e-3
    assert bool((m8[0] == 0).all())


def _opt_model(layers=2, hidden=256, heads=4, ffn=512, vocab=512, width=6):
    from llm_mixed_q_b200.models.opt_quantized import OPTQuantizedConfig, OPTQuantizedForCausalLM

    cfg = OPTQuantizedConfig(hidden_size=hidden, num_hidden_layers=layers, ffn_dim=ffn, num_attention_heads=heads,
                             vocab_size=vocab, max_position_embeddings=256, quant_config={"default": bfp_cfg(width)})
    torch.manual_seed(0)
    model = OPTQuantizedForCausalLM(cfg).eval()
    with torch.no_grad():          # non-trivial biases / LN parameters so every fused term is exercised
        for n, p in model.named_parameters():
            if n.endswith(".bias"):
                p.normal_(0, 0.02)
            elif "layer_norm.weight" in n:
                p.add_(0.1 * torch.randn_like(p))
    return model.cuda()


@pytest.mark.parametrize("width", [6, 4])
def test_fused_opt_layer_matches_op_by_op_and_oracle(width):
    """Whole model through the 8-kernel fused layers vs (a) the same modules op by op and (b) the oracle's restatement of
    the reference forward (torch emulation on the same device).  Statistical tolerance as in test_gpu_models.py."""
    model = _opt_model(width=width)
    dec = model.model.decoder
    g = torch.Generator(device="cuda").manual_seed(1)
    ids = torch.randint(0, 512, (3, 128), device="cuda", generator=g)
    with torch.no_grad():
        # snapshot BEFORE the first forward: PTQ overwrites the parameters, and block_fp is not idempotent (a block max that
        # rounds down onto a power of two gets a smaller shared exponent the second time) — the reference quantises once
        sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
        dec.fused_glue, dec.fused_attention = False, False
        ref = model(input_ids=ids, labels=ids)                 # also performs the PTQ weight overwrite
        dec.fused_glue, dec.fused_attention = True, True
        assert dec.layers[0]._fused_plan(128) is not None
        out = model(input_ids=ids, labels=ids)
        qc = model.config.quant_config
        o_logits, o_loss = opt_ref.opt_forward(sd, qc, ids, num_layers=2, num_heads=4, labels=ids)
    spread = float(ref.logits.std())
    for name, r_logits, r_loss in (("op-by-op", ref.logits, float(ref.loss)), ("oracle", o_logits, float(o_loss))):
        err = (out.logits - r_logits).abs()
        assert abs(float(out.loss) - r_loss) <= 2e-3 * abs(r_loss), (name, float(out.loss), r_loss)
        assert float(err.mean()) <= 0.02 * spread and float(err.max()) <= 0.5 * spread, (name, float(err.mean()), float(err.max()), spread)


def test_graphed_forward_replays_the_eager_result():
    """llm_mixed_q_b200.utils.graphs.GraphedForward: the whole fused forward captured in one CUDA graph (every kernel of the C ABI is
    a plain stream launch) reproduces the eager forward bit for bit, for new token ids on every replay."""
    from llm_mixed_q_b200.utils.graphs import GraphedForward

    model = _opt_model()
    g = torch.Generator(device="cuda").manual_seed(7)
    ids = [torch.randint(0, 512, (3, 128), device="cuda", generator=g) for _ in range(3)]
    with torch.no_grad():
        eager = [model(input_ids=i, labels=i) for i in ids]
    runner = GraphedForward(model, 3, 128)
    assert runner.graph is not None, runner.error
    for i, e in zip(ids, eager):
        loss = runner(i.cpu().pin_memory())                    # host ids: the H2D copy is part of the call
        assert torch.equal(loss, e.loss) and torch.equal(runner.logits, e.logits)
    # a model whose forward needs a host-side decision (padding mask -> op-by-op path) still works through the eager fallback
    assert float(runner(ids[0])) == float(eager[0].loss)


def test_programmatic_dependent_launch_does_not_change_results():
    """bq_set_pdl: the GEMM / attention / norm+quantize kernels start their set-up under the tail of the previous kernel and wait
    (griddepcontrol.wait) before touching global memory.  Results must be bit-identical with the switch off, eagerly and from a
    captured graph, and stable over repeated back-to-back forwards (a kernel reading its predecessor's output too early would
    show up as run-to-run differences)."""
    from llm_mixed_q_b200 import _lib as L
    from llm_mixed_q_b200.utils.graphs import GraphedForward

    lib = L.load()
    before = lib.bq_get_pdl()
    model = _opt_model()
    g = torch.Generator(device="cuda").manual_seed(9)
    ids = torch.randint(0, 512, (3, 128), device="cuda", generator=g)
    try:
        lib.bq_set_pdl(0)
        with torch.no_grad():
            want = model(input_ids=ids, labels=ids)
        lib.bq_set_pdl(1)
        with torch.no_grad():
            for _ in range(20):
                got = model(input_ids=ids, labels=ids)
                assert torch.equal(got.logits, want.logits) and torch.equal(got.loss, want.loss)
        runner = GraphedForward(model, 3, 128)
        assert runner.graph is not None, runner.error
        for _ in range(20):
            loss = runner(ids)
            assert torch.equal(loss, want.loss) and torch.equal(runner.logits, want.logits)
    finally:
        lib.bq_set_pdl(before)


def test_graphed_forward_llama():
    """Same for the fused Llama layers (RMSNorm+quantize, RoPE+quantize, attention, SiLU*up+quantize)."""
    import json
    import os

    from conftest import GOLD
    from llm_mixed_q_b200.models.llama_quantized import LlamaQuantizedConfig, LlamaQuantizedForCausalLM
    from llm_mixed_q_b200.utils.graphs import GraphedForward

    with open(os.path.join(GOLD, "configs.json")) as f:
        qc = json.load(f)["raw"]["bfp_6bit.toml"]
    torch.manual_seed(0)
    cfg = LlamaQuantizedConfig(hidden_size=128, intermediate_size=352, num_hidden_layers=2, num_attention_heads=2, vocab_size=512,
                               max_position_embeddings=128, quant_config=qc)
    model = LlamaQuantizedForCausalLM(cfg).eval().cuda()
    g = torch.Generator(device="cuda").manual_seed(8)
    ids = [torch.randint(0, 512, (2, 128), device="cuda", generator=g) for _ in range(2)]
    with torch.no_grad():
        eager = [model(input_ids=i, labels=i) for i in ids]
        assert model.model.layers[0]._fused_plan(128) is not None
    runner = GraphedForward(model, 2, 128)
    assert runner.graph is not None, runner.error
    for i, e in zip(ids, eager):
        loss = runner(i)
        assert torch.equal(loss, e.loss) and torch.equal(runner.logits, e.logits)


def test_fused_layer_falls_back_when_not_eligible():
    model = _opt_model()
    dec = model.model.decoder
    assert dec.layers[0]._fused_plan(100) is None              # S % 16 != 0: k^T blocks would straddle rows
    ids = torch.randint(0, 512, (2, 100), device="cuda")
    with torch.no_grad():
        out = model(input_ids=ids, labels=ids)
    assert torch.isfinite(out.loss)


@pytest.mark.parametrize("tomlname,init", [("bfp_6bit.toml", 0.02), ("block_minifloat.toml", 1.5)])
def test_fused_llama_layer_matches_op_by_op(tomlname, init):
    """Llama through the fused layer (RMSNorm+quantize, prequantised GEMMs with residual / v-quantiser epilogues, token-major
    RoPE, one attention kernel with the post-matmul 1/sqrt(d), SiLU*up quantised for down_proj) vs the same modules op by op."""
    import json
    import os

    from conftest import GOLD
    from llm_mixed_q_b200.models.llama_quantized import LlamaQuantizedConfig, LlamaQuantizedForCausalLM

    with open(os.path.join(GOLD, "configs.json")) as f:
        qc = json.load(f)["raw"][tomlname]
    torch.manual_seed(0)
    cfg = LlamaQuantizedConfig(hidden_size=128, intermediate_size=352, num_hidden_layers=2, num_attention_heads=2, vocab_size=512,
                               max_position_embeddings=128, initializer_range=init, quant_config=qc)
    model = LlamaQuantizedForCausalLM(cfg).eval().cuda()
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "layernorm.weight" in n:
                p.add_(0.1 * torch.randn_like(p))
    ids = torch.randint(0, 512, (3, 128), device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    with torch.no_grad():
        model.model.fused_glue = False
        ref = model(input_ids=ids, labels=ids)                 # also performs the PTQ weight overwrite
        model.model.fused_glue = True
        assert model.model.layers[0]._fused_plan(128) is not None
        out = model(input_ids=ids, labels=ids)
        # a padded batch must take the op-by-op path and still agree with itself
        am = torch.ones_like(ids)
        am[0, :5] = 0
        padded = model(input_ids=ids, attention_mask=am, labels=ids)
    assert torch.isfinite(padded.loss)
    spread = float(ref.logits.std())
    err = (out.logits - ref.logits).abs()
    assert abs(float(out.loss) - float(ref.loss)) <= 5e-3 * abs(float(ref.loss)), (float(out.loss), float(ref.loss))
    assert float(err.mean()) <= 0.03 * spread and float(err.max()) <= 0.75 * spread, (float(err.mean()), float(err.max()), spread)


@pytest.mark.parametrize("kind", ["block_fp", "block_minifloat"])
@pytest.mark.parametrize("heads,d", [(4, 64), (2, 128)])
def test_rope_quantize_kernel_is_bit_identical_to_rope_ops_plus_quantizers(kind, heads, d):
    """bq_rope_quantize == apply_rotary_pos_emb (torch ops on the quantised tables) followed by the q / k^T quantizers of matmul_0."""
    from llm_mixed_q_b200.models.llama_quantized.modeling_llama import LlamaRotaryEmbedding
    from llm_mixed_q_b200.models.quantize.quantized_functions.attention import quantize_qkv
    from llm_mixed_q_b200.models.quantize.quantized_functions.rotary_positional_encoding import (apply_token_major,
                                                                                                  apply_token_major_quantized)

    B, S, H = 3, 96, heads * d
    if kind == "block_fp":
        m0 = {"name": "block_fp", "bypass": False}
        for p in ("data_in", "weight"):
            m0.update({f"{p}_width": 6, f"{p}_exponent_width": 8, f"{p}_exponent_bias": 127, f"{p}_block_size": [1, 16]})
    else:
        m0 = {"name": "block_minifloat", "bypass": False}
        for p in ("data_in", "weight"):
            m0.update({f"{p}_width": 8, f"{p}_exponent_width": 4, f"{p}_exponent_bias_width": 8, f"{p}_block_size": [1, 16]})
    rope_cfg = {"name": "integer", "bypass": False, "data_in_width": 8, "data_in_frac_width": 7}
    g = torch.Generator(device="cuda").manual_seed(heads * d)
    q = torch.randn(B, S, H, device="cuda", generator=g) * 2
    k = torch.randn(B, S, H, device="cuda", generator=g) * 2
    q[0, 3] = 0
    k[1, 16:32, :d] = 0                                     # all-zero k^T blocks
    k[2, 5, 7] = 1e-9                                       # pass-through element
    rot = LlamaRotaryEmbedding(d, max_position_embeddings=256).cuda()
    cos, sin = rot(q, seq_len=S)
    for pos in (torch.arange(S, device="cuda")[None].expand(B, S), torch.randint(0, S, (B, S), device="cuda", generator=g), None):
        pid = pos if pos is not None else torch.arange(S, device="cuda")[None]
        q4, k4 = apply_token_major(q.view(B, S, heads, d), k.view(B, S, heads, d), cos, sin, pid, rope_cfg)
        Qr, Kr, _ = quantize_qkv(q4.reshape(B, S, H), k4.reshape(B, S, H), None, m0, m0, heads)
        got = apply_token_major_quantized(q, k, cos, sin, pid if pos is not None else None, rope_cfg, m0, heads)
        assert got is not None
        assert torch.equal(got[0].view(torch.int16), Qr.reshape(B, S, H).view(torch.int16))
        assert torch.equal(got[1].view(torch.int16), Kr.reshape(B, S, H).view(torch.int16))
    # not eligible: block sizes other than [1,16] -> None, the caller falls back
    m_odd = dict(m0, data_in_block_size=[1, 32])
    assert apply_token_major_quantized(q, k, cos, sin, None, rope_cfg, m_odd, heads) is None


@pytest.mark.parametrize("stream", [1, 0])
@pytest.mark.parametrize("kind", ["block_fp", "block_minifloat"])
def test_silu_mul_quantize_is_bit_identical_to_torch_silu_mul_then_quantizer(kind, stream):
    """Q(silu(gate) * up) in one kernel == torch's silu, multiply, then the quantizer (the A/B switch must not change the result:
    the silu*mul prologue lives in the per-slot kernel — a bulk-copy streaming variant holding both tiles in shared memory was
    measured slower, 0.316 vs 0.188 ms at 4096 x 11008: one CTA of 8 warps per SM cannot hide expf + IEEE division)."""
    from llm_mixed_q_b200 import _lib as L
    from llm_mixed_q_b200.models.quantize.quantized_functions.fused_glue import silu_mul_quantize
    from llm_mixed_q_b200.models.quantize.quantizers import block_fp_quantizer, block_minifloat_quantizer

    if kind == "block_fp":
        fmt, qz = ("block_fp", dict(width=6, exponent_width=8, exponent_bias=127)), lambda t: block_fp_quantizer(t, 6, 8, 127, [1, 16], True)
    else:
        fmt = ("block_minifloat", dict(width=8, exponent_width=4, exponent_bias_width=8))
        qz = lambda t: block_minifloat_quantizer(t, 8, 4, 8, [1, 16], True)
    g = torch.Generator(device="cuda").manual_seed(3)
    L.load().bq_set_stream_quantizer(stream)
    try:
        for rows, I in [(300, 11008), (64, 352), (5, 16), (1030, 1024)]:
            gate = torch.randn(rows, I, device="cuda", generator=g) * 3
            up = torch.randn(rows, I, device="cuda", generator=g)
            gate[0, :16] = 0
            up[1, 16:32] = 0
            gate[2, 5] = -100.0                          # silu underflows to -0
            want = qz(torch.nn.functional.silu(gate) * up)
            got = silu_mul_quantize(gate, up, fmt)
            assert_bf16_carrier_equal(got, want, torch.nn.functional.silu(gate) * up)
            got32 = silu_mul_quantize(gate, up, fmt, out_dtype=torch.float32)
            assert torch.equal(got32, want)
    finally:
        L.load().bq_set_stream_quantizer(1)


@pytest.mark.parametrize("kind", ["block_fp", "block_minifloat", "block_log"])
def test_gated_silu_gemm_epilogue_is_bit_identical_to_two_gemms_and_the_silu_mul_quantizer(kind):
    """gated_silu_prequantized (one GEMM over the interleaved gate / up weights, silu * up and down_proj's x-quantizer in the
    epilogue: bq_gemm_bf16_tn_ex act = 2) against gate GEMM + up GEMM + silu_mul_quantize — the reference's
    down_proj(act_fn(gate_proj(x)) * up_proj(x)) operand (modeling_llama.py:84).  Pair tiles, single-CTA tiles, ragged rows."""
    import copy

    from llm_mixed_q_b200 import _lib as L
    from llm_mixed_q_b200.models.quantize import get_quantized_cls
    from llm_mixed_q_b200.models.quantize.quantized_functions.fused_glue import silu_mul_quantize
    from llm_mixed_q_b200.models.quantize.quantized_modules import linear as QL

    if kind == "block_fp":
        fmt = ("block_fp", dict(width=6, exponent_width=8, exponent_bias=127))
    elif kind == "block_minifloat":
        fmt = ("block_minifloat", dict(width=8, exponent_width=4, exponent_bias_width=8))
    else:
        fmt = ("block_log", dict(width=8, exponent_bias_width=8))
    cfg = bfp_cfg(6)
    g = torch.Generator(device="cuda").manual_seed(11)
    n0 = L.launch_counts().get("gemm_bf16_tn_kernel<epilogue>", 0)
    for rows, K, I in [(300, 256, 352), (4096, 512, 1408), (64, 128, 64), (130, 64, 11008)]:
        gate = get_quantized_cls("linear", cfg)(K, I, bias=False, config=copy.deepcopy(cfg)).cuda()
        up = get_quantized_cls("linear", cfg)(K, I, bias=False, config=copy.deepcopy(cfg)).cuda()
        with torch.no_grad():
            gate.weight.mul_(20.0)                      # gate pre-activations of a few units: both tails of the SiLU
            up.weight[:16].zero_()                      # an all-zero block of products
        assert QL.gated_silu_fusable(gate, up)
        x = torch.randn(rows, K, device="cuda", generator=g)
        xq = QL.quantize_operand_bf16(x, "block_fp", dict(width=6, exponent_width=8, exponent_bias=127), [1, 16], True)
        gv, uv = gate.forward_prequantized(xq), up.forward_prequantized(xq)
        want = silu_mul_quantize(gv, uv, fmt)
        got = QL.gated_silu_prequantized(gate, up, xq, fmt)
        assert got.shape == want.shape and got.dtype == torch.bfloat16
        if kind != "block_log":
            assert torch.equal(got.view(torch.int16), want.view(torch.int16)), (kind, rows, K, I)
        else:
            # block_log in a 16-bit carrier (include/bq.h, "carrier rule"): bit-identical to the exact fp32 quantizer output wherever
            # that is >= 2^-126 in magnitude; an all-zero block stays 0 and smaller outputs come out as 0 or 2^-126
            want32 = silu_mul_quantize(gv, uv, fmt, out_dtype=torch.float32)
            prod = torch.nn.functional.silu(gv) * uv
            zero_blocks = (prod.view(rows, I // 16, 16) == 0).all(-1, keepdim=True).expand(rows, I // 16, 16).reshape(rows, I)
            want32 = torch.where(zero_blocks, torch.zeros_like(want32), want32)
            tiny = 2.0 ** -126
            big = want32.abs() >= tiny
            assert torch.equal(got.float()[big], want32[big]), (kind, rows, K, I)
            small = got.float()[~big].abs()
            assert bool(((small == 0) | (small == tiny)).all())
    assert L.launch_counts().get("gemm_bf16_tn_kernel<epilogue>", 0) - n0 >= 4
    # argument contract of act = 2 (include/bq.h)
    lib = L.load()
    from llm_mixed_q_b200.models.quantize.quantizers.utils import make_format

    A = torch.zeros(32, 64, device="cuda", dtype=torch.bfloat16)
    B = torch.zeros(256, 64, device="cuda", dtype=torch.bfloat16)
    C = torch.zeros(32, 128, device="cuda", dtype=torch.bfloat16)
    f = make_format("block_fp", b0=1, b1=16, width=6, exponent_width=8, exponent_bias=127)
    ep = L.BqGemmEpilogue()
    ep.scale, ep.act, ep.out_dtype = 1.0, 2, L.BQ_BF16
    assert lib.bq_gemm_bf16_tn_ex(A.data_ptr(), B.data_ptr(), C.data_ptr(), ctypes.byref(ep), 32, 256, 64, 64, 64, 128, L.stream_ptr()) == 2
    ep.qfmt = ctypes.pointer(f)
    assert lib.bq_gemm_bf16_tn_ex(A.data_ptr(), B.data_ptr(), C.data_ptr(), ctypes.byref(ep), 32, 256, 64, 64, 64, 64, L.stream_ptr()) == 1
    ep.out_dtype = L.BQ_F32
    assert lib.bq_gemm_bf16_tn_ex(A.data_ptr(), B.data_ptr(), C.data_ptr(), ctypes.byref(ep), 32, 256, 64, 64, 64, 128, L.stream_ptr()) == 2


@pytest.mark.parametrize("d", [128, 64])
@pytest.mark.parametrize("kind", ["block_fp", "block_minifloat"])
def test_rope_gemm_epilogue_is_bit_identical_to_gemm_then_rope_quantize(kind, d):
    """bq_gemm_bf16_tn_rope (q_proj / k_proj with the rotary embedding and matmul_0's operand quantizer in the GEMM epilogue) against
    the fp32 GEMM followed by bq_rope_quantize — which is itself bit-identical to the reference's torch ops + quantizers
    (test above).  Pair tiles and single-CTA tiles, default and explicit positions, both block directions."""
    import copy

    from llm_mixed_q_b200.models.llama_quantized.modeling_llama import LlamaRotaryEmbedding
    from llm_mixed_q_b200.models.quantize import get_quantized_cls
    from llm_mixed_q_b200.models.quantize.quantized_functions.rotary_positional_encoding import (apply_token_major_quantized,
                                                                                                 rope_quantize_operands)
    from llm_mixed_q_b200.models.quantize.quantized_modules import linear as QL

    if kind == "block_fp":
        m0 = {"name": "block_fp", "bypass": False}
        for p in ("data_in", "weight"):
            m0.update({f"{p}_width": 6, f"{p}_exponent_width": 8, f"{p}_exponent_bias": 127, f"{p}_block_size": [1, 16]})
    else:
        m0 = {"name": "block_minifloat", "bypass": False}
        for p in ("data_in", "weight"):
            m0.update({f"{p}_width": 8, f"{p}_exponent_width": 4, f"{p}_exponent_bias_width": 8, f"{p}_block_size": [1, 16]})
    rope_cfg = {"name": "integer", "bypass": False, "data_in_width": 8, "data_in_frac_width": 7}
    cfg = bfp_cfg(6)
    g = torch.Generator(device="cuda").manual_seed(13 + d)
    for B, S, heads, K in [(2, 128, 4, 256), (1, 2048, 2, 128), (3, 48, 2, 64), (2, 2048, 8, 512)]:
        H = heads * d
        lin_q = get_quantized_cls("linear", cfg)(K, H, bias=False, config=copy.deepcopy(cfg)).cuda()
        lin_k = get_quantized_cls("linear", cfg)(K, H, bias=False, config=copy.deepcopy(cfg)).cuda()
        with torch.no_grad():
            lin_q.weight.mul_(8.0)
            lin_k.weight.mul_(8.0)
            lin_k.weight[5].zero_()                             # a feature whose k^T blocks are all zero
        assert QL.rope_epilogue_fusable(lin_q, d) and QL.rope_epilogue_fusable(lin_k, d)
        x = torch.randn(B * S, K, device="cuda", generator=g)
        xq = QL.quantize_operand_bf16(x, "block_fp", dict(width=6, exponent_width=8, exponent_bias=127), [1, 16], True)
        q32, k32 = lin_q.forward_prequantized(xq), lin_k.forward_prequantized(xq)
        rot = LlamaRotaryEmbedding(d, max_position_embeddings=max(S, 256)).cuda()
        cos, sin = rot(q32, seq_len=S)
        for pos in (None, torch.randint(0, S, (B, S), device="cuda", generator=g)):
            want = apply_token_major_quantized(q32.view(B, S, H), k32.view(B, S, H), cos, sin, pos, rope_cfg, m0, heads)
            assert want is not None
            cos_t, sin_t, p64, fq, fk = rope_quantize_operands(cos, sin, pos, rope_cfg, m0, B, S, d)
            got_q = QL.rope_prequantized(lin_q, xq, cos_t, sin_t, p64, fq, S, d, False)
            got_k = QL.rope_prequantized(lin_k, xq, cos_t, sin_t, p64, fk, S, d, True)
            assert torch.equal(got_q.view(torch.int16), want[0].reshape(B * S, H).view(torch.int16)), (kind, d, B, S, pos is None)
            assert torch.equal(got_k.view(torch.int16), want[1].reshape(B * S, H).view(torch.int16)), (kind, d, B, S, pos is None)
            # q | k | v in one launch over the concatenated weights (bq_gemm_bf16_tn_qkv_rope): same bits as the separate launches
            if H % 256 == 0 and (B * S) % 16 == 0:
                lin_v = get_quantized_cls("linear", cfg)(K, H, bias=False, config=copy.deepcopy(cfg)).cuda()
                with torch.no_grad():
                    lin_v.weight.mul_(8.0)
                v_fmt = ("block_fp", dict(width=6, exponent_width=8, exponent_bias=127)) if kind == "block_fp" else \
                        ("block_minifloat", dict(width=8, exponent_width=4, exponent_bias_width=8))
                assert QL.qkv_rope_fusable(lin_q, lin_k, lin_v, d)
                want_v = lin_v.forward_prequantized(xq, out_format=v_fmt)
                q1, k1, v1 = QL.qkv_rope_prequantized(lin_q, lin_k, lin_v, xq, cos_t, sin_t, p64, fq, fk, v_fmt, S, d)
                assert torch.equal(q1.view(torch.int16), got_q.view(torch.int16)) and torch.equal(k1.view(torch.int16), got_k.view(torch.int16))
                assert torch.equal(v1.view(torch.int16), want_v.view(torch.int16))


def test_llama_layer_with_gated_epilogue_equals_three_launch_mlp():
    """The fused Llama layer with the gated GEMM epilogue is bit-identical to the same layer with gate GEMM + up GEMM + silu*mul
    quantizer (QL.GATED_EPILOGUE = False), block_minifloat W4A4-style config and block_log (split plan)."""
    import json
    import os

    from conftest import GOLD
    from llm_mixed_q_b200.models.llama_quantized import LlamaQuantizedConfig, LlamaQuantizedForCausalLM
    from llm_mixed_q_b200.models.quantize.quantized_modules import linear as QL

    with open(os.path.join(GOLD, "configs.json")) as f:
        raw = json.load(f)["raw"]
    for name in ("block_minifloat.toml", "block_log.toml", "bfp_6bit.toml"):
        if name not in raw:
            continue
        cfg = LlamaQuantizedConfig(hidden_size=256, intermediate_size=352, num_hidden_layers=2, num_attention_heads=2, vocab_size=512,
                                   max_position_embeddings=128, quant_config=raw[name])
        torch.manual_seed(0)
        model = LlamaQuantizedForCausalLM(cfg).eval().cuda()
        ids = torch.randint(0, 512, (2, 128), device="cuda")
        assert model.model.layers[0]._fused_plan(128) is not None
        outs = []
        try:
            for on in (True, False):
                # (RoPE in the q / k GEMM epilogues and q | k | v as one launch: same A/B, same requirement)
                QL.GATED_EPILOGUE = QL.ROPE_EPILOGUE = QL.QKV_ONE_LAUNCH = on
                with torch.no_grad():
                    outs.append(model(ids).logits)
        finally:
            QL.GATED_EPILOGUE = QL.ROPE_EPILOGUE = QL.QKV_ONE_LAUNCH = True
        assert torch.equal(outs[0], outs[1]), name
        assert getattr(model.model.layers[0].mlp.gate_proj, "_gu_cache", None) is not None
        if name != "block_log.toml":
            assert getattr(model.model.layers[0].self_attn.q_proj, "_qkv_cache", None) is not None


def test_fused_llama_block_log_takes_the_split_attention_plan():
    import json
    import os

    from conftest import GOLD
    from llm_mixed_q_b200.models.llama_quantized import LlamaQuantizedConfig, LlamaQuantizedForCausalLM

    with open(os.path.join(GOLD, "configs.json")) as f:
        qc = json.load(f)["raw"]["block_log.toml"]
    cfg = LlamaQuantizedConfig(hidden_size=128, intermediate_size=352, num_hidden_layers=1, num_attention_heads=2, vocab_size=512,
                               max_position_embeddings=128, quant_config=qc)
    model = LlamaQuantizedForCausalLM(cfg).eval().cuda()
    plan = model.model.layers[0]._fused_plan(128)
    # matmul_0 / matmul_1 keep k / v in fp32 (reference matmul.py:286-297): three-kernel attention, fused glue around it
    assert plan is not None and plan["mode"] == "split" and plan["q_in"][0] == "block_log"
    cfg2 = LlamaQuantizedConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=1, num_attention_heads=4, vocab_size=512,
                                max_position_embeddings=128, quant_config=qc)
    assert LlamaQuantizedForCausalLM(cfg2).eval().cuda().model.layers[0]._fused_plan(128) is None       # head_dim 16: op by op
