"""GPU: the two north-star variants of the quantized Linear (csrc/gemm_xform_sm100.cu) against the default two-launch path.

  bq_linear_fused      x-quantizer inside the GEMM prologue   (reference quantized_modules/linear.py:59-76 in ONE launch)
  bq_pack_weight /     weights as w + 0.5 bits per element (the reference's cost model, quantized_layer_profiler.py:18-27),
  bq_gemm_packed_tn    decoded to bf16 in the mainloop

Both use the same quantiser arithmetic and the same K order of fp32 accumulation as bq_linear, so results are compared BIT FOR BIT."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cfg(xw=6, ww=6):
    cfg = {"name": "block_fp", "bypass": False, "is_ptq": True}
    for p, w in (("data_in", xw), ("weight", ww), ("bias", ww)):
        cfg.update({f"{p}_width": w, f"{p}_exponent_width": 8, f"{p}_exponent_bias": 127,
                    f"{p}_block_size": [16] if p == "bias" else [1, 16]})
    return cfg


def _linear(K, N, cfg, seed=0, std=0.02):
    from llm_mixed_q_b200.models.quantize import get_quantized_cls

    torch.manual_seed(seed)
    lin = get_quantized_cls("linear", cfg)(K, N, bias=True, config=cfg).cuda().eval()
    with torch.no_grad():
        lin.weight.normal_(0, std)
        lin.bias.normal_(0, 0.02)
    return lin


@pytest.mark.parametrize("width", [2, 3, 4, 5, 6, 7, 8])
def test_packed_weights_round_trip_exactly(width):
    """pack -> GEMM against an identity activation = unpack: every weight comes back bit for bit (products with 1.0 and sums with
    zeros are exact), for every field width; bits per element = width + 0.5."""
    from llm_mixed_q_b200.models.quantize.quantized_modules.linear import gemm_packed, pack_weight
    from llm_mixed_q_b200.models.quantize.quantizers import block_fp_quantizer

    K, N = 512, 96
    g = torch.Generator(device="cuda").manual_seed(width)
    w = torch.randn(N, K, device="cuda", generator=g) * torch.logspace(-3, 1, N, device="cuda")[:, None]
    w[5, 32:48] = 0.0                                                   # an all-zero block
    wq = block_fp_quantizer(w, width, 8, 127, [1, 16], False)
    packed, bad = pack_weight(wq, width, 8, 127)
    assert packed.shape == (N, K // 256 * (32 * width + 16)) and packed.numel() * 8 == N * K * (width + 0.5)
    assert bad == 0
    eye = torch.eye(K, device="cuda", dtype=torch.bfloat16)
    back = gemm_packed(eye, packed, width, 8, 127, N).t().contiguous()   # [N, K]
    assert torch.equal(back.view(torch.int32), (wq + 0.0).view(torch.int32))


def test_pack_counts_pass_through_elements():
    """|x| <= 1e-8 is returned UNQUANTISED by the reference (block_fp.py:93-94): such an element is not on the block's grid, the
    packed form rounds it (to 0 here) and reports it."""
    from llm_mixed_q_b200.models.quantize.quantized_modules.linear import pack_weight
    from llm_mixed_q_b200.models.quantize.quantizers import block_fp_quantizer

    w = torch.randn(32, 256, device="cuda") * 0.02
    w[3, 7] = 3e-9
    w[9, 100] = -7e-9
    wq = block_fp_quantizer(w, 6, 8, 127, [1, 16], False)
    assert float(wq[3, 7]) == pytest.approx(3e-9) and float(wq[9, 100]) == pytest.approx(-7e-9)
    _, bad = pack_weight(wq, 6, 8, 127)
    assert bad == 2


@pytest.mark.parametrize("M,K,N,ww", [(64, 512, 256, 6), (16, 1024, 96, 4), (512, 512, 384, 6), (300, 768, 256, 3), (1024, 2048, 2048, 6),
                                      (8, 256, 128, 5), (32, 512, 384, 6), (100, 1024, 2048, 2), (128, 4096, 1024, 8), (1, 256, 256, 7)])
def test_packed_linear_is_bit_identical_to_the_bf16_cache_path(M, K, N, ww):
    """LinearBlockFP with PACKED_WEIGHTS: same bits as the default path (bf16 weight cache) for decode-sized M (BN = 32 tiles) and
    prefill-sized M (BN = 128 tiles), ragged M, several widths."""
    from llm_mixed_q_b200.models.quantize.quantized_modules import linear as lin_mod

    cfg = _cfg(6, ww)
    lin = _linear(K, N, cfg)
    x = torch.randn(M, K, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    with torch.no_grad():
        ref = lin(x)
        lin_mod.PACKED_WEIGHTS = True
        try:
            out = lin(x)
            bits = lin.packed_bits_per_element()
        finally:
            lin_mod.PACKED_WEIGHTS = False
    assert bits == ww + 0.5
    bad = lin._wq_packed[2]
    if bad == 0:
        assert torch.equal(ref.view(torch.int32), out.view(torch.int32))
    else:
        # pass-through weights (|w| <= 1e-8, ~4e-7 of N(0, 0.02) draws) are not on the block grid: the packed form holds 0 for them,
        # the bf16 cache their rounded value — each moves an output by at most 1e-8 * |x| (+ one rounding of the sum)
        assert bad <= 1e-5 * N * K
        tol = bad * 1e-8 * float(x.abs().max()) + 2.0 ** -22 * float(ref.abs().max())
        assert float((ref - out).abs().max()) <= tol


@pytest.mark.parametrize("M,K,N", [(128, 64, 32), (300, 512, 256), (4096, 1024, 512), (77, 2048, 96)])
@pytest.mark.parametrize("kind", ["block_fp", "block_minifloat"])
def test_fused_prologue_linear_is_bit_identical_to_two_launch_path(M, K, N, kind):
    """x-quantizer in the GEMM prologue (one launch) vs quantize kernel + GEMM (two launches): same bits, incl. ragged M, inputs with
    zero blocks, huge and tiny magnitudes (slow quantiser paths) and a 3-D activation."""
    from llm_mixed_q_b200.models.quantize import get_quantized_cls
    from llm_mixed_q_b200.models.quantize.quantized_modules import linear as lin_mod

    if kind == "block_fp":
        cfg = _cfg(6, 6)
    else:
        cfg = {"name": "block_minifloat", "bypass": False, "is_ptq": True}
        for p in ("data_in", "weight", "bias"):
            cfg.update({f"{p}_width": 8, f"{p}_exponent_width": 4, f"{p}_exponent_bias_width": 8,
                        f"{p}_block_size": [16] if p == "bias" else [1, 16]})
    torch.manual_seed(3)
    lin = get_quantized_cls("linear", cfg)(K, N, bias=True, config=cfg).cuda().eval()
    with torch.no_grad():
        lin.weight.normal_(0, 1.0 if kind == "block_minifloat" else 0.02)
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn(M, K, device="cuda", generator=g) * 3
    x[0, :16] = 0.0
    x[min(5, M - 1), 16:32] *= 1e30
    x[min(7, M - 1), 32:48] *= 1e-30
    x[min(9, M - 1), 48] = 4e-9                                         # pass-through element
    with torch.no_grad():
        ref = lin(x)
        lin_mod.FUSED_PROLOGUE = True
        try:
            out = lin(x)
            out3 = lin(x.view(1, M, K)) if M % 1 == 0 else None
        finally:
            lin_mod.FUSED_PROLOGUE = False
    assert torch.equal(ref.view(torch.int32), out.view(torch.int32))
    assert torch.equal(ref.view(torch.int32), out3.view(M, N).view(torch.int32))


def test_variant_entry_points_reject_bad_arguments():
    import ctypes

    from llm_mixed_q_b200 import _lib as L
    from llm_mixed_q_b200.models.quantize.quantizers.utils import make_format

    lib = L.load()
    f6 = make_format("block_fp", width=6, exponent_width=8, exponent_bias=127, b0=1, b1=16)
    f12 = make_format("block_fp", width=12, exponent_width=8, exponent_bias=127, b0=1, b1=16)
    assert lib.bq_packed_weight_bytes(ctypes.byref(f6), 64, 512) == 64 * 2 * (32 * 6 + 16)
    assert lib.bq_packed_weight_bytes(ctypes.byref(f12), 64, 512) == 0          # > 8 bits: not packable
    assert lib.bq_packed_weight_bytes(ctypes.byref(f6), 64, 500) == 0           # K % 256
    x = torch.zeros(128, 96, device="cuda")
    w = torch.zeros(32, 96, device="cuda", dtype=torch.bfloat16)
    y = torch.zeros(128, 32, device="cuda")
    assert lib.bq_linear_fused(ctypes.byref(f6), x.data_ptr(), 128, 96, 96, w.data_ptr(), 32, None, y.data_ptr(), 32, None) != 0   # K % 64
