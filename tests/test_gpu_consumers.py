"""GPU parity of the quantizer consumers: tcgen05 GEMM, quantized Linear modules, quantized matmul/bmm.

Tolerance (stated): operands are bit-identical quantised values; products of two <=8-bit-significand values are
exact in fp32, so the only difference to the reference's fp32 GEMM is accumulation order.  For a length-K
dot product with fp32 accumulation  |err| <= gamma * sum_k |a_k b_k|,  gamma ~ sqrt(K) * 2^-24 typically and
K * 2^-24 worst case.  We assert  |out - exact| <= 4 * sqrt(K) * 2^-24 * (|A| @ |B|) + pass_through_term  against an
fp64 evaluation of the same quantised operands, and the same bound (x2, both sides round) against the golden
fp32 outputs of the reference.
pass_through_term: elements with |x| <= 1e-8 are returned UNQUANTISED by the reference (block_fp.py:93-94); they are
arbitrary fp32 values and the bf16 operand carrier rounds them (relative 2^-9): each contributes at most
2^-9 * 1e-8 * |b| — an absolute 2e-11 per term, stated here and in DESIGN.md."""
import copy

import pytest
import torch

from conftest import bits_equal, f32
from oracle import oracle as O

pytestmark = pytest.mark.gpu
EPS = 2.0 ** -24


def assert_gemm_close(out, exact64, absprod64, K, factor=4.0, b_abs_colsum=None):
    bound = factor * (K ** 0.5) * EPS * absprod64 + 1e-30
    if b_abs_colsum is not None:
        bound = bound + (2.0 ** -9) * 1e-8 * b_abs_colsum
    err = (out.double() - exact64).abs()
    worst = float((err / bound).max())
    assert worst <= 1.0, f"GEMM error {worst:.2f}x the stated fp32-accumulation-order bound"


def test_gemm_kernel_vs_fp64():
    import ctypes

    from llm_mixed_q_b200 import _lib as L

    lib = L.load()
    g = torch.Generator(device="cuda").manual_seed(0)
    for (b, M, N, K) in [(1, 128, 256, 64), (1, 4096, 4096, 4096), (1, 1000, 520, 328), (3, 200, 136, 64), (4, 512, 64, 512),
                         (2, 300, 100, 1000), (1, 1, 8, 8), (1, 129, 257, 72), (5, 2048, 2048, 64), (2, 2048, 64, 2048),
                         (1, 4864, 5200, 4096)]:       # last: B > 40 MB -> banded raster, bands of 16 + 3 row blocks, ragged N
        A = torch.randn(b, M, K, device="cuda", generator=g).to(torch.bfloat16)
        B = torch.randn(b, N, K, device="cuda", generator=g).to(torch.bfloat16)
        bias = torch.randn(N, device="cuda", generator=g)
        C = torch.full((b, M, N), float("nan"), device="cuda")
        rc = lib.bq_gemm_bf16_tn(A.data_ptr(), B.data_ptr(), C.data_ptr(), bias.data_ptr(), b, M, N, K, K, K, N, M * K, N * K,
                                 M * N, L.stream_ptr())
        L.check(rc, "gemm")
        exact = A.double() @ B.double().transpose(1, 2) + bias.double()
        absprod = A.double().abs() @ B.double().abs().transpose(1, 2) + bias.double().abs()
        assert_gemm_close(C, exact, absprod, K)


def _module_for(cfg, in_f, out_f):
    from llm_mixed_q_b200.models.quantize import get_quantized_cls

    name = cfg["name"] if not cfg.get("bypass") else "block_fp"
    return get_quantized_cls("linear", {"name": name})(in_f, out_f, bias=True, config=copy.deepcopy(cfg)).cuda()


def test_linear_modules_vs_reference_goldens(golden_consumers):
    arrays, cases = golden_consumers
    for c in [c for c in cases if c["op"] == "linear"]:
        k = c["key"]
        x, w, b = f32(arrays[k + "_x"]).cuda(), f32(arrays[k + "_w"]), f32(arrays[k + "_b"])
        lin = _module_for(c["config"], w.shape[1], w.shape[0])
        with torch.no_grad():
            lin.weight.copy_(w)
            lin.bias.copy_(b)
        y = lin(x)
        ref = f32(arrays[k + "_y"]).reshape(y.shape)
        # the PTQ in-place overwrite leaves bit-identical parameters (reference linear.py:66-70)
        assert bits_equal(lin.weight.detach().cpu(), f32(arrays[k + "_wq"])), (k, c["config_name"])
        assert bits_equal(lin.bias.detach().cpu(), f32(arrays[k + "_bq"])), (k, c["config_name"])
        if not c["config"].get("bypass"):
            assert lin.weight_requires_quantisation is False
        xq = O.operand_quantizer(c["config"], "data_in", True)(x.cpu()).double() if not c["config"].get("bypass") else x.cpu().double()
        wq, bq = lin.weight.detach().cpu().double(), lin.bias.detach().cpu().double()
        absprod = xq.abs() @ wq.abs().T + bq.abs()
        assert_gemm_close(y.cpu(), ref.double(), absprod, w.shape[1], factor=8.0)
        y2 = lin(x)                                    # steady state: same result again
        assert torch.equal(y, y2)


def test_matmul_functions_vs_reference_goldens(golden_consumers):
    from llm_mixed_q_b200.models.quantize import get_quantized_func

    arrays, cases = golden_consumers
    for c in [c for c in cases if c["op"] in ("bmm", "matmul")]:
        k = c["key"]
        x, yb = f32(arrays[k + "_x"]).cuda(), f32(arrays[k + "_y"]).cuda()
        yv = yb.transpose(-1, -2) if c["y_transposed"] else yb
        fn = get_quantized_func(c["op"], c["config"])
        out = fn(x, yv, config=copy.deepcopy(c["config"]))
        ref = f32(arrays[k + "_o"]).reshape(out.shape)
        K = x.shape[-1]
        scale = float(ref.abs().max())
        assert float((out.cpu() - ref).abs().max()) <= 8 * (K ** 0.5) * EPS * max(scale, 1.0) * 16, (k, c["config_name"])


CFG_BFP6 = {"name": "block_fp", "bypass": False, "is_ptq": True}
for _p in ("data_in", "weight", "bias"):
    CFG_BFP6.update({f"{_p}_width": 6, f"{_p}_exponent_width": 8, f"{_p}_exponent_bias": 127,
                     f"{_p}_block_size": [16] if _p == "bias" else [1, 16]})


@pytest.mark.parametrize("width", [6, 4])
def test_linear_config2_vs_device_oracle(width):
    """BASELINE config 2: M=K=N=4096, block_fp W6A6 / W4A4."""
    from llm_mixed_q_b200.models.quantize import get_quantized_cls

    cfg = copy.deepcopy(CFG_BFP6)
    for p in ("data_in", "weight", "bias"):
        cfg[f"{p}_width"] = width
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(4096, 4096, device="cuda", generator=g)
    w = torch.randn(4096, 4096, device="cuda", generator=g) * 0.02
    b = torch.randn(4096, device="cuda", generator=g) * 0.02
    lin = get_quantized_cls("linear", cfg)(4096, 4096, bias=True, config=cfg).cuda()
    with torch.no_grad():
        lin.weight.copy_(w)
        lin.bias.copy_(b)
    y = lin(x)
    yo, wq, bq = O.linear_forward(x, w, b, cfg)          # oracle on the same device
    assert bits_equal(lin.weight.detach(), wq) and bits_equal(lin.bias.detach(), bq)
    xq = O.operand_quantizer(cfg, "data_in", True)(x).double()
    exact = xq @ wq.double().T + bq.double()
    absprod = xq.abs() @ wq.double().abs().T + bq.double().abs()
    assert_gemm_close(y, exact, absprod, 4096)
    assert_gemm_close(yo, exact, absprod, 4096)           # the reference's own fp32 GEMM obeys the same bound


def test_attention_bmms_vs_device_oracle():
    """bmm_0 (q @ k^T view, y blocked along key positions) and bmm_1 (P @ v) at OPT-1.3B head geometry."""
    from llm_mixed_q_b200.models.quantize import get_quantized_func

    g = torch.Generator(device="cuda").manual_seed(2)
    BH, S, d = 16, 2048, 64
    q = torch.randn(BH, S, d, device="cuda", generator=g)
    k = torch.randn(BH, S, d, device="cuda", generator=g)
    v = torch.randn(BH, S, d, device="cuda", generator=g)
    fn = get_quantized_func("bmm", CFG_BFP6)
    s = fn(q, k.transpose(1, 2), config=CFG_BFP6)
    xq = O.operand_quantizer(CFG_BFP6, "data_in", True)(q).double()
    yq = O.operand_quantizer(CFG_BFP6, "weight", True)(k.transpose(1, 2)).double()
    assert_gemm_close(s, xq @ yq, xq.abs() @ yq.abs(), d)
    mask = torch.triu(torch.ones(S, S, dtype=torch.bool, device="cuda"), diagonal=1)
    p = torch.softmax(s.masked_fill(mask, torch.finfo(torch.float32).min), dim=-1)
    o = fn(p, v, config=CFG_BFP6)
    pq = O.operand_quantizer(CFG_BFP6, "data_in", True)(p).double()
    vq = O.operand_quantizer(CFG_BFP6, "weight", True)(v).double()
    assert_gemm_close(o, pq @ vq, pq.abs() @ vq.abs(), S, b_abs_colsum=vq.abs().sum(dim=1, keepdim=True))


def test_block_log_matmul_leaves_y_unquantised():
    from llm_mixed_q_b200.models.quantize import get_quantized_func

    cfg = {"name": "block_log", "bypass": False, "data_in_width": 8, "data_in_exponent_bias_width": 8,
           "data_in_block_size": [1, 16], "weight_width": 8, "weight_exponent_bias_width": 8, "weight_block_size": [1, 16]}
    g = torch.Generator(device="cuda").manual_seed(4)
    x = torch.randn(3, 64, 48, device="cuda", generator=g)
    y = torch.randn(3, 48, 32, device="cuda", generator=g)
    out = get_quantized_func("bmm", cfg)(x, y, config=cfg)
    ref = torch.bmm(O.block_log_quantize(x, 8, 8, [1, 16], True), y)     # reference matmul.py:293-296
    torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-5)
    assert get_quantized_func("bmm", {"name": "log"}) is get_quantized_func("bmm", cfg)


@pytest.mark.parametrize("name", ["block_log", "block_fp12"])
def test_general_route_matmul_runs_on_the_split_tensor_core_gemm(name):
    """Operands that are not bf16-exact (block_log: y unquantised fp32; formats wider than 8 significant bits) go through the
    batched fp16-plane split GEMM: same fp32-accumulation-order bound as the fp32 matmul of the reference, k^T views, ragged
    shapes, 2-D y, and no SIMT library GEMM launched."""
    from llm_mixed_q_b200 import _lib as L
    from llm_mixed_q_b200.models.quantize import get_quantized_func

    if name == "block_log":
        cfg = {"name": "block_log", "bypass": False, "data_in_width": 4, "data_in_exponent_bias_width": 8,
               "data_in_block_size": [1, 16], "weight_width": 4, "weight_exponent_bias_width": 8, "weight_block_size": [1, 16]}
        qx = lambda t, multi: O.block_log_quantize(t, 4, 8, [1, 16], multi)
        qy = lambda t, multi: t
    else:
        cfg = {"name": "block_fp", "bypass": False}
        for p in ("data_in", "weight"):
            cfg.update({f"{p}_width": 12, f"{p}_exponent_width": 8, f"{p}_exponent_bias": 127, f"{p}_block_size": [1, 16]})
        qx = qy = lambda t, multi: O.block_fp_quantize(t, 12, 8, 127, [1, 16], multi)
    g = torch.Generator(device="cuda").manual_seed(11)
    n0 = L.launch_counts()["gemm_bf16_tn_kernel<split>"]
    cases = [("bmm", (8, 512, 128), (8, 128, 512), True), ("bmm", (8, 512, 512), (8, 512, 128), False),
             ("bmm", (3, 300, 1000), (3, 1000, 100), False), ("matmul", (2, 5, 64, 48), (2, 5, 48, 40), False),
             ("matmul", (4, 200, 64), (64, 72), False)]
    for style, xs, ys, kt in cases:
        x = torch.randn(*xs, device="cuda", generator=g)
        if kt:
            y = torch.randn(ys[0], ys[2], ys[1], device="cuda", generator=g).transpose(1, 2)      # k^T view
        else:
            y = torch.randn(*ys, device="cuda", generator=g)
        x[..., 3, :] *= 1e-5
        x[..., 4, :] = 0
        out = get_quantized_func(style, cfg)(x, y, config=copy.deepcopy(cfg))
        f3 = lambda t: torch.flatten(t, 0, -3) if t.ndim > 2 else t                   # reference matmul.py:166-190: flatten, quantise, restore
        xq = qx(f3(x).contiguous(), x.ndim > 2).reshape(x.shape)
        yq = qy(f3(y).contiguous(), y.ndim > 2).reshape(y.shape)
        exact = xq.double() @ yq.double()
        absprod = xq.double().abs() @ yq.double().abs()
        assert out.shape == exact.shape
        assert_gemm_close(out, exact, absprod, xs[-1])
        assert_gemm_close(torch.matmul(xq, yq), exact, absprod, xs[-1])            # the reference's own route obeys the same bound
    assert L.launch_counts()["gemm_bf16_tn_kernel<split>"] - n0 == len(cases)


def _oracle_attention(q, k, v, cfg, heads, score_div=1.0):
    """op-by-op reference composition (modeling_opt.py:237-323) with the oracle's quantizers, on the same device."""
    B, S, H = q.shape
    d = H // heads

    def shape(t):
        return t.view(B, S, heads, d).transpose(1, 2).contiguous().view(B * heads, S, d)

    q3, k3, v3 = shape(q), shape(k), shape(v)
    s = O.matmul_forward(q3, k3.transpose(1, 2), cfg, style="bmm") / score_div
    mask = torch.triu(torch.full((S, S), torch.finfo(torch.float32).min, device=q.device), diagonal=1)
    s = torch.max(s + mask, torch.tensor(torch.finfo(torch.float32).min, device=q.device))
    p = torch.softmax(s, dim=-1)
    o = O.matmul_forward(p, v3, cfg, style="bmm")
    return o.view(B, heads, S, d).transpose(1, 2).reshape(B, S, H), p


@pytest.fixture(params=[0, 1], ids=["fast_exp", "precise_exp"])
def attention_exp_mode(request):
    """Both numerator modes of the fused attention kernel (bq_set_attention_precise_exp)."""
    from llm_mixed_q_b200 import _lib as L

    L.load().bq_set_attention_precise_exp(request.param)
    yield request.param
    L.load().bq_set_attention_precise_exp(0)


@pytest.fixture(params=[1, 0], ids=["dual_pipeline_p_in_tmem", "single_pipeline_p_in_smem"])
def attention_pipeline(request):
    """head_dim 64 runs on the two-pipeline kernel with P in tensor memory by default; the single-pipeline kernel (what head_dim
    128 uses) stays selectable (bq_set_attention_dual_pipeline) and must give the same answers."""
    from llm_mixed_q_b200 import _lib as L

    L.load().bq_set_attention_dual_pipeline(request.param)
    yield request.param
    L.load().bq_set_attention_dual_pipeline(1)


def test_fused_attention_pipelines_agree_bit_for_bit():
    """Same arithmetic in both kernels (only the tiling of the key axis differs: 64- vs 128-key tiles change the order in which the
    row sum is accumulated) — the quantised probabilities, and therefore the
    outputs, must agree almost everywhere; rows whose sums differ in the last ulp may flip single probabilities by one step."""
    from llm_mixed_q_b200 import _lib as L
    from llm_mixed_q_b200.models.quantize.quantized_functions.attention import fused_causal_attention

    g = torch.Generator(device="cuda").manual_seed(5)
    B, heads, d, S = 2, 4, 64, 1024
    q = torch.randn(B, S, heads * d, device="cuda", generator=g) * 0.5
    k = torch.randn(B, S, heads * d, device="cuda", generator=g)
    v = torch.randn(B, S, heads * d, device="cuda", generator=g)
    outs = []
    try:
        for dual in (1, 0):
            L.load().bq_set_attention_dual_pipeline(dual)
            outs.append(fused_causal_attention(q, k, v, CFG_BFP6, CFG_BFP6, heads))
    finally:
        L.load().bq_set_attention_dual_pipeline(1)
    diff = (outs[0] - outs[1]).abs()
    assert float((diff > 0).float().mean()) <= 2e-2, float((diff > 0).float().mean())
    assert float(diff.max()) <= (2.0 ** -5) * float(v.abs().max())


@pytest.mark.parametrize("S", [64, 128, 200, 1024, 2048])
def test_fused_causal_attention_vs_op_by_op_oracle(S, attention_exp_mode, attention_pipeline):
    """The fused kernel computes the same function as bmm_0 -> mask -> softmax -> bmm_1.  Scores agree to fp32
    accumulation order; the row sum of the softmax is accumulated in a different order from 2-ulp ex2.approx
    exponentials, so a probability can differ by a few ulp BEFORE quantisation and, when it sits on a rounding boundary,
    by one quantisation step after it (SURVEY.md hard part 6).  Stated tolerance: every output within ONE quantisation
    step of the largest possible probability (2^-5 for W6) times max|v|; at most 1e-4 of the outputs off by more than
    1 % of max|v|; mean error <= 2e-4."""
    from llm_mixed_q_b200.models.quantize.quantized_functions.attention import fusable, fused_causal_attention

    g = torch.Generator(device="cuda").manual_seed(100 + S)
    B, heads, d = 2, 4, 64
    H = heads * d
    q = torch.randn(B, S, H, device="cuda", generator=g) * 0.5
    k = torch.randn(B, S, H, device="cuda", generator=g)
    v = torch.randn(B, S, H, device="cuda", generator=g)
    assert fusable(CFG_BFP6, CFG_BFP6, d, S)
    out = fused_causal_attention(q, k, v, CFG_BFP6, CFG_BFP6, heads)
    ref, p = _oracle_attention(q, k, v, CFG_BFP6, heads)
    err = (out - ref).abs()
    vmax = float(v.abs().max())
    assert float(err.max()) <= (2.0 ** -5) * vmax, float(err.max())
    assert float((err > 0.01 * vmax).float().mean()) <= 1e-4
    assert float(err.mean()) <= 2e-4, float(err.mean())
    # row 0 attends to a single key: p = 1 -> Q(1) = 31/32 exactly, out = 31/32 * Q(v[0])
    vq = O.operand_quantizer(CFG_BFP6, "weight", True)(v.view(B, S, heads, d).transpose(1, 2).reshape(B * heads, S, d))
    exp0 = (vq[:, 0, :] * (31.0 / 32.0)).view(B, heads, d).reshape(B, H)
    assert torch.equal(out[:, 0, :], exp0)


@pytest.mark.parametrize("q_scale", [1.0, 2.5])
def test_fused_attention_peaked_softmax_pass_through_tiers(q_scale, attention_exp_mode, attention_pipeline):
    """Peaked softmax rows (what trained checkpoints produce): most probabilities are <= 1e-8, which the reference returns
    UNQUANTISED (block_fp.py:93-94).  The kernel serves such blocks from two in-line tiers — the whole block below the threshold
    (truncated to the bf16 carrier) and mixed blocks (fp32 quantise + per-element select) — and must still be the op-by-op
    composition within the tolerance of test_fused_causal_attention_vs_op_by_op_oracle."""
    from llm_mixed_q_b200.models.quantize.quantized_functions.attention import fused_causal_attention

    g = torch.Generator(device="cuda").manual_seed(77)
    B, heads, d, S = 2, 4, 64, 1024
    H = heads * d
    q = torch.randn(B, S, H, device="cuda", generator=g) * q_scale
    k = torch.randn(B, S, H, device="cuda", generator=g)
    v = torch.randn(B, S, H, device="cuda", generator=g)
    out = fused_causal_attention(q, k, v, CFG_BFP6, CFG_BFP6, heads)
    ref, p = _oracle_attention(q, k, v, CFG_BFP6, heads)
    visible = torch.tril(torch.ones(S, S, dtype=torch.bool, device="cuda"))
    tiny = (p <= 1e-8) & visible
    blocks = (p * visible).view(B * heads, S, S // 16, 16)
    all_tiny = ((blocks <= 1e-8).all(-1) & (blocks > 0).any(-1)).float().mean()
    mixed = ((blocks <= 1e-8).any(-1) & (blocks > 1e-8).any(-1)).float().mean()
    assert float(tiny.float().sum() / visible.sum() / (B * heads)) > 0.5          # the tiers ARE what this input exercises
    assert float(all_tiny) > 0.02 and float(mixed) > 0.05, (float(all_tiny), float(mixed))     # shares of ALL S x S / 16 blocks, half of them masked
    err = (out - ref).abs()
    vmax = float(v.abs().max())
    assert float(err.max()) <= (2.0 ** -5) * vmax, float(err.max())
    assert float((err > 0.01 * vmax).float().mean()) <= 1e-4
    assert float(err.mean()) <= 2e-4, float(err.mean())


def _oracle_attention_masked(q, k, v, cfg, heads, valid, causal, score_div=1.0):
    """op-by-op reference composition with an additive key-padding mask: decoder (causal + padding, clamped at finfo.min —
    opt_quantized/modeling_opt.py:520-548, :266-270) or encoder (padding only, no clamp — bert_quantized/modeling_bert.py:366-435)."""
    B, S, H = q.shape
    d = H // heads
    neg = torch.finfo(torch.float32).min

    def shape(t):
        return t.view(B, S, heads, d).transpose(1, 2).contiguous().view(B * heads, S, d)

    q3, k3, v3 = shape(q), shape(k), shape(v)
    s = (O.matmul_forward(q3, k3.transpose(1, 2), cfg, style="bmm") / score_div).view(B, heads, S, S)
    pad = torch.zeros(B, 1, 1, S, device=q.device).masked_fill(~valid[:, None, None, :], neg)
    if causal:
        mask = (torch.triu(torch.full((S, S), neg, device=q.device), diagonal=1)[None, None] + pad).clamp(min=neg)
        s = torch.max(s + mask, torch.tensor(neg, device=q.device))
    else:
        s = s + pad
    p = torch.softmax(s, dim=-1).view(B * heads, S, S)
    o = O.matmul_forward(p, v3, cfg, style="bmm")
    return o.view(B, heads, S, d).transpose(1, 2).reshape(B, S, H)


@pytest.mark.parametrize("S,d,causal,pattern", [(200, 64, True, "right"), (1024, 64, True, "right"), (640, 128, True, "right"),
                                                (384, 64, True, "holes"), (80, 64, False, "right"), (512, 64, False, "right"),
                                                (333, 64, False, "holes"), (384, 128, False, "right"), (512, 64, False, "none")])
def test_fused_attention_with_key_padding_and_bidirectional_masks(S, d, causal, pattern, attention_pipeline):
    """bq_attention_masked: the key-padding bitmap (padded OPT / Llama batches) and the bidirectional mode (BERT) against the
    op-by-op composition; same stated tolerance as the causal test.  "holes": arbitrary key subsets (key 0 kept, so every causal
    row has a key) — some rows meet 32-key slices with no visible key before their first valid one."""
    import math

    from llm_mixed_q_b200.models.quantize.quantized_functions.attention import fused_causal_attention, key_mask_bits

    g = torch.Generator(device="cuda").manual_seed(S + d + (7 if causal else 0))
    B, heads = 3, 2
    H = heads * d
    q = torch.randn(B, S, H, device="cuda", generator=g) * 0.5
    k = torch.randn(B, S, H, device="cuda", generator=g)
    v = torch.randn(B, S, H, device="cuda", generator=g)
    valid = torch.ones(B, S, dtype=torch.bool, device="cuda")
    if pattern == "right":
        valid[1, S // 2 + 5:] = False
        valid[2, 33:] = False
    elif pattern == "holes":
        valid[1] = torch.rand(S, device="cuda", generator=g) > 0.6
        valid[2, 1:130] = False
        valid[:, 0] = True
    sd = math.sqrt(d) if not causal else 1.0
    out = fused_causal_attention(q, k, v, CFG_BFP6, CFG_BFP6, heads, score_div=sd, causal=causal, key_mask=key_mask_bits(valid))
    ref = _oracle_attention_masked(q, k, v, CFG_BFP6, heads, valid, causal, score_div=sd)
    assert torch.isfinite(out).all()
    err = (out - ref).abs()
    vmax = float(v.abs().max())
    assert float(err.max()) <= (2.0 ** -5) * vmax, float(err.max())
    assert float((err > 0.01 * vmax).float().mean()) <= 1e-4
    assert float(err.mean()) <= 2e-4, float(err.mean())
    # the x-quantised output form agrees with the quantizer applied to the fp32 form
    from llm_mixed_q_b200.models.quantize.quantizers import block_fp_quantizer
    oq = fused_causal_attention(q, k, v, CFG_BFP6, CFG_BFP6, heads, score_div=sd, causal=causal, key_mask=key_mask_bits(valid),
                                out_cfg=CFG_BFP6)
    want = block_fp_quantizer(out, 6, 8, 127, [1, 16], True)
    big = out.abs() > 1e-8
    assert torch.equal(oq.float()[big], want[big])


def test_key_mask_bits_layout_and_argument_checks():
    from llm_mixed_q_b200.models.quantize.quantized_functions.attention import fused_causal_attention, key_mask_bits

    valid = torch.zeros(2, 200, dtype=torch.bool, device="cuda")
    valid[0, [0, 31, 32, 199]] = True
    valid[1, :] = True
    bits = key_mask_bits(valid)
    assert bits.shape == (2, 8) and bits.dtype == torch.int32
    w = bits.cpu().numpy().astype("int64") & 0xFFFFFFFF
    assert w[0, 0] == (1 | (1 << 31)) and w[0, 1] == 1 and w[0, 6] == (1 << 7) and w[0, 7] == 0
    assert w[1, 6] == 0xFF and w[1, 5] == 0xFFFFFFFF and w[1, 7] == 0          # keys >= S cleared
    q = torch.randn(2, 200, 128, device="cuda")
    with pytest.raises(ValueError):
        fused_causal_attention(q, q, q, CFG_BFP6, CFG_BFP6, 2, causal=False, key_mask=bits[:, :4].contiguous().to(torch.int64))
    with pytest.raises(ValueError):                         # bitmap shorter than the key tiles (BQ_ERR_BAD_ARG)
        fused_causal_attention(q, q, q, CFG_BFP6, CFG_BFP6, 2, causal=False, key_mask=bits[:, :4].contiguous())


def test_fused_causal_attention_head_dim_128_and_score_div(attention_exp_mode):
    """Llama-7B geometry: d = 128, scores divided by sqrt(d) after matmul_0 (modeling_llama.py:309-314).  torch-CUDA
    evaluates `tensor / python_float` as a multiplication by the fp32 reciprocal; the kernel does the same."""
    import math

    from llm_mixed_q_b200.models.quantize.quantized_functions.attention import fusable, fused_causal_attention

    g = torch.Generator(device="cuda").manual_seed(77)
    B, heads, d, S = 2, 3, 128, 640
    H = heads * d
    q = torch.randn(B, S, H, device="cuda", generator=g)
    k = torch.randn(B, S, H, device="cuda", generator=g)
    v = torch.randn(B, S, H, device="cuda", generator=g)
    assert fusable(CFG_BFP6, CFG_BFP6, d, S)
    out = fused_causal_attention(q, k, v, CFG_BFP6, CFG_BFP6, heads, score_div=math.sqrt(d))
    ref, _ = _oracle_attention(q, k, v, CFG_BFP6, heads, score_div=math.sqrt(d))
    err = (out - ref).abs()
    vmax = float(v.abs().max())
    assert float(err.max()) <= (2.0 ** -5) * vmax, float(err.max())
    assert float((err > 0.01 * vmax).float().mean()) <= 1e-4
    assert float(err.mean()) <= 2e-4, float(err.mean())


@pytest.mark.parametrize("d", [64, 128])
def test_fused_attention_quantised_output_equals_quantizer_of_fp32_output(d):
    """bq_attention_causal_q == x-quantizer(out_proj) applied to bq_attention_causal's fp32 result, bit for bit
    (the epilogue quantises the same TMEM accumulator the fp32 variant stores)."""
    from llm_mixed_q_b200.models.quantize.quantized_functions.attention import fused_causal_attention, output_quantizable
    from llm_mixed_q_b200.models.quantize.quantizers import block_fp_quantizer

    g = torch.Generator(device="cuda").manual_seed(5 + d)
    B, heads, S = 2, 2, 384
    H = heads * d
    q = torch.randn(B, S, H, device="cuda", generator=g)
    k = torch.randn(B, S, H, device="cuda", generator=g)
    v = torch.randn(B, S, H, device="cuda", generator=g)
    assert output_quantizable(CFG_BFP6, H)
    o32 = fused_causal_attention(q, k, v, CFG_BFP6, CFG_BFP6, heads)
    oq = fused_causal_attention(q, k, v, CFG_BFP6, CFG_BFP6, heads, out_cfg=CFG_BFP6)
    assert oq.dtype == torch.bfloat16
    want = block_fp_quantizer(o32, 6, 8, 127, [1, 16], True)
    passthrough = o32.abs() <= 1e-8          # returned UNQUANTISED by the reference (block_fp.py:93-94): the bf16 carrier rounds them
    assert torch.equal(oq.float()[~passthrough], want[~passthrough])
    assert torch.equal(oq[passthrough], want[passthrough].to(torch.bfloat16))


@pytest.mark.parametrize("mode", ["f16x2", "bf16x3"])
def test_fp32_equivalent_linear_for_unquantised_layers(mode):
    """lm_head-style fp32 Linear through the split tensor-core GEMMs: error must stay inside the same
    fp32-accumulation-order bound an fp32 GEMM obeys (the reference's F.linear is checked against it too)."""
    from llm_mixed_q_b200.models.quantize.quantized_functions.fp32_linear import fp32_linear, split2_rows, split3

    g = torch.Generator(device="cuda").manual_seed(9)
    x = torch.randn(1024, 2048, device="cuda", generator=g)
    x[5] *= 1e-6                                    # rows of very different magnitude: the per-row scale must absorb it
    x[6] *= 3e4
    x[7] = 0
    w = torch.randn(5000, 2048, device="cuda", generator=g) * 0.02
    bias = torch.randn(5000, device="cuda", generator=g)
    planes = split3(x).float()
    assert float((planes.sum(0).double() - x.double()).abs().max()) <= 2.0 ** -24 * float(x.abs().max())
    p2, inv = split2_rows(x)
    rec = (p2[0].double() + p2[1].double()) * inv.double()[:, None]
    rowmax = x.abs().amax(1, keepdim=True).double()
    assert bool(((rec - x.double()).abs() <= 2.0 ** -21 * rowmax + 1e-300).all())
    y = fp32_linear(x, w, bias, mode=mode)
    exact = x.double() @ w.double().T + bias.double()
    absprod = x.double().abs() @ w.double().abs().T + bias.double().abs()
    assert_gemm_close(y, exact, absprod, 2048)
    assert_gemm_close(torch.nn.functional.linear(x, w, bias), exact, absprod, 2048)
    rel = float(((y.double() - exact).abs() / absprod.clamp_min(1e-300)).max())
    assert rel < 1e-6, rel
