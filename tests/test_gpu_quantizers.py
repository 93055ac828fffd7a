"""GPU parity: the sm_100a quantizer kernels, called through the reference-facing Python API (which goes
through the C ABI), are BIT-EXACT against (a) the committed golden vectors produced by the unmodified
reference and (b) the oracle restatement evaluated on the same device at BASELINE sizes."""
import hashlib

import pytest
import torch

from conftest import bits_equal, case_input, case_kwargs, f32, n_bits_diff
from golden_inputs import hashed_input
from oracle import oracle as O

pytestmark = pytest.mark.gpu
BLOCKED = ("block_fp", "block_minifloat", "block_log")


def product(name):
    from llm_mixed_q_b200.models.quantize.quantizers import QUANTIZER_MAP

    return QUANTIZER_MAP[name]


def run_product(case, x):
    fn = product(case["fmt"])
    kw = case_kwargs(case)
    if case["fmt"] in BLOCKED:
        return fn(x, block_size=list(case["block_size"]), skip_first_dim=case["skip_first_dim"], **kw)
    return fn(x, **kw)


def test_all_golden_vectors_bit_exact(golden_quantizers):
    arrays, cases = golden_quantizers
    bad = []
    for case in cases:
        x = case_input(arrays, case).cuda()
        x_before = x.clone()
        y = run_product(case, x)
        assert y.dtype == torch.float32 and y.is_cuda and tuple(y.shape) == tuple(x.shape)
        assert bits_equal(x, x_before), "input must be untouched"
        exp = f32(arrays[case["key"]]).reshape(y.shape)
        if not bits_equal(y.cpu(), exp):
            bad.append((case["key"], case["fmt"], case["kwargs"], case["layout"], case["input"], case["block_size"],
                        n_bits_diff(y.cpu(), exp)))
    assert not bad, f"{len(bad)}/{len(cases)} golden mismatches, first: {bad[:8]}"


def test_hashed_large_cases_bit_exact(golden_hashed):
    for case in golden_hashed:
        x = hashed_input(case).cuda()
        fn = product(case["fmt"])
        if case["block_size"] is not None:
            y = fn(x, block_size=case["block_size"], skip_first_dim=case["skip_first_dim"], **case["kwargs"])
        else:
            y = fn(x, **case["kwargs"])
        h = hashlib.sha256(y.cpu().contiguous().view(torch.int32).numpy().tobytes()).hexdigest()
        assert h == case["sha256"], (case["tag"], case["fmt"], case["kwargs"])


FORMATS = [
    ("block_fp", dict(width=6, exponent_width=8, exponent_bias=127)),
    ("block_fp", dict(width=4, exponent_width=8, exponent_bias=None)),
    ("block_minifloat", dict(width=8, exponent_width=4, exponent_bias_width=8)),
    ("block_minifloat", dict(width=4, exponent_width=2, exponent_bias_width=8)),
    ("block_log", dict(width=8, exponent_bias_width=8)),
    ("block_log", dict(width=4, exponent_bias_width=8)),
]


def both(name, kw, x, block_size, skip):
    y = product(name)(x, block_size=block_size, skip_first_dim=skip, **kw)
    yo = O.QUANTIZERS[name](x, block_size=block_size, skip_first_dim=skip, **kw)
    return y, yo


@pytest.mark.parametrize("name,kw", FORMATS)
@pytest.mark.parametrize("sigma", [1e-3, 0.02, 1.0, 30.0])
def test_baseline_shapes_vs_device_oracle(name, kw, sigma):
    """OPT-1.3B activation shape [8, 2048, 2048] and the config-2 weight shape [4096, 4096]."""
    g = torch.Generator(device="cuda").manual_seed(1234)
    x = torch.randn(8, 2048, 2048, device="cuda", generator=g) * sigma
    x.view(-1)[::97] = 0
    x[:, ::7, 64:96] = 0                  # all-zero blocks
    y, yo = both(name, kw, x, [1, 16], True)
    assert n_bits_diff(y, yo) == 0
    del y, yo, x
    w = torch.randn(4096, 4096, device="cuda", generator=g) * sigma
    y, yo = both(name, kw, w, [1, 16], False)
    assert n_bits_diff(y, yo) == 0


@pytest.mark.parametrize("name,kw", FORMATS)
def test_softmax_probs_with_masked_zero_blocks(name, kw):
    """P tensor: post-softmax rows under a causal mask (exact zeros, all-zero blocks, tensor-global min)."""
    g = torch.Generator(device="cuda").manual_seed(7)
    s = torch.randn(16, 1024, 1024, device="cuda", generator=g) * 3
    mask = torch.triu(torch.ones(1024, 1024, dtype=torch.bool, device="cuda"), diagonal=1)
    p = torch.softmax(s.masked_fill(mask, torch.finfo(torch.float32).min), dim=-1)
    y, yo = both(name, kw, p, [1, 16], True)
    assert n_bits_diff(y, yo) == 0


@pytest.mark.parametrize("name,kw", FORMATS)
def test_log2_cliffs_vs_device_oracle(name, kw):
    """block maxima / elements at 2^k (1 + d 2^-23) and sqrt(2) 2^k (1 + d 2^-23): the ceil/floor/round(log2f) cliffs."""
    k = torch.arange(-40, 41, device="cuda", dtype=torch.float32)
    d = torch.arange(-8, 17, device="cuda", dtype=torch.int32)
    for base in (1.0, 2.0 ** 0.5):
        b = torch.tensor(base, dtype=torch.float32, device="cuda").view(torch.int32)
        mant = (b + d).view(torch.float32)                       # [25]
        vals = (mant[None, :] * torch.exp2(k)[:, None]).reshape(-1)   # [81*25]
        x = torch.zeros(vals.numel(), 16, device="cuda")
        x[:, 0] = vals
        x[:, 1:] = vals[:, None] * torch.linspace(-0.9, 0.9, 15, device="cuda")[None, :]
        y, yo = both(name, kw, x, [1, 16], True)
        assert n_bits_diff(y, yo) == 0


@pytest.mark.parametrize("bs", [[1, 4], [1, 8], [1, 32], [1, 64], [1, 128], [2, 16], [16, 16], [3, 5], [16], [1, 4096]])
def test_other_block_shapes_vs_device_oracle(bs):
    g = torch.Generator(device="cuda").manual_seed(3)
    for shape, skip in [((5, 33, 200), True), ((130, 264), False), ((130, 264), True), ((520,), False)]:
        x = torch.randn(*shape, device="cuda", generator=g)
        for name, kw in FORMATS[0:1] + FORMATS[2:3] + FORMATS[4:5]:
            y, yo = both(name, kw, x, bs, skip)
            assert n_bits_diff(y, yo) == 0, (shape, name, bs)


def test_strided_and_transposed_views():
    g = torch.Generator(device="cuda").manual_seed(5)
    base = torch.randn(6, 300, 64, device="cuda", generator=g)
    kT = base.transpose(1, 2)                                    # k^T of bmm_0: blocks along the strided dim
    for name, kw in FORMATS:
        y, yo = both(name, kw, kT, [1, 16], True)
        assert y.is_contiguous() and n_bits_diff(y, yo) == 0, name
    rows = torch.randn(64, 512, device="cuda", generator=g)[:, 128:384]     # row-strided slice
    for name, kw in FORMATS:
        y, yo = both(name, kw, rows, [1, 16], True)
        assert n_bits_diff(y, yo) == 0, name


def test_elementwise_formats_vs_device_oracle():
    from llm_mixed_q_b200.models.quantize.quantizers import (integer_quantizer, minifloat_denorm_quantizer,
                                                             minifloat_ieee_quantizer)

    g = torch.Generator(device="cuda").manual_seed(11)
    for sigma in (1e-3, 1.0, 300.0):
        x = torch.randn(1000, 4096, device="cuda", generator=g) * sigma
        x.view(-1)[::5] = 0
        for kw in (dict(width=8, exponent_width=4, exponent_bias=None), dict(width=8, exponent_width=4, exponent_bias=7),
                   dict(width=4, exponent_width=2, exponent_bias=None)):
            assert n_bits_diff(minifloat_denorm_quantizer(x, **kw), O.minifloat_denorm_quantize(x, **kw)) == 0
            assert n_bits_diff(minifloat_ieee_quantizer(x, **kw), O.minifloat_ieee_quantize(x, **kw)) == 0
        assert n_bits_diff(integer_quantizer(x, 8, 7), O.integer_quantize(x, 8, 7)) == 0
        xo = x[:, 1:4001:3]                                      # non-contiguous element-wise input
        assert n_bits_diff(minifloat_denorm_quantizer(xo, 8, 4, 7), O.minifloat_denorm_quantize(xo, 8, 4, 7)) == 0


def test_properties_at_full_size():
    """size-independent properties at the P-tensor scale of config 3 (a [64, 2048, 2048] slice)."""
    from llm_mixed_q_b200.models.quantize.quantizers import block_fp_quantizer

    g = torch.Generator(device="cuda").manual_seed(21)
    x = torch.randn(64, 2048, 2048, device="cuda", generator=g)
    y = block_fp_quantizer(x, 6, 8, 127, [1, 16], True)
    # odd symmetry (exact: |x| + 1e-9 and sign(x + 1e-9) are symmetric once |x| > 1e-8)
    assert torch.equal(block_fp_quantizer(-x, 6, 8, 127, [1, 16], True), -y)
    # blocks are independent and order-free: reversing the 16 elements of every block commutes with quantisation
    xr = x.view(-1, 16).flip(1).reshape(x.shape)
    assert torch.equal(block_fp_quantizer(xr, 6, 8, 127, [1, 16], True).view(-1, 16).flip(1).reshape(x.shape), y)
    # every block lies on one exponent grid: y = q * 2^(E-5), |q| <= 31 integer, E = ceil(log2(block max))
    step = torch.exp2(torch.ceil(torch.log2(x.view(-1, 16).abs().amax(1, keepdim=True))) - 5)
    q = y.view(-1, 16) / step
    assert torch.equal(q, q.round()) and float(q.abs().max()) <= 31
    # quantisation error is at most half a step (+ the 1e-9 epsilon), except for the saturated block maximum (one step)
    assert float(((y - x).view(-1, 16).abs() / step).max()) <= 1.0 + 1e-6
    del xr, q
    # pure function: same input -> same bits
    assert n_bits_diff(block_fp_quantizer(x, 6, 8, 127, [1, 16], True), y) == 0


def test_empty_and_tiny_inputs():
    from llm_mixed_q_b200.models.quantize.quantizers import block_fp_quantizer, block_log_quantizer

    assert block_fp_quantizer(torch.zeros(0, 16, device="cuda"), 6, 8, 127, [1, 16], True).shape == (0, 16)
    z = torch.zeros(4, 32, device="cuda")
    assert n_bits_diff(block_fp_quantizer(z, 6, 8, 127, [1, 16], True), O.block_fp_quantize(z, 6, 8, 127, [1, 16], True)) == 0
    assert n_bits_diff(block_log_quantizer(z, 8, 8, [1, 16], True), O.block_log_quantize(z, 8, 8, [1, 16], True)) == 0
    one = torch.tensor([0.3], device="cuda")
    assert n_bits_diff(block_fp_quantizer(one, 6, 8, 127, [16], False), O.block_fp_quantize(one, 6, 8, 127, [16], False)) == 0


def test_ste_backward():
    from llm_mixed_q_b200.models.quantize.quantizers import block_fp_quantizer

    x = torch.randn(4, 64, device="cuda", requires_grad=True)
    y = block_fp_quantizer(x, 6, 8, 127, [1, 16], True)
    y.sum().backward()
    assert torch.equal(x.grad, torch.ones_like(x))


def test_stream_kernel_matches_rows_kernel_and_oracle():
    """The bulk-copy streaming kernel (dense, block 16 / element-wise) against the per-slot rows kernel and the oracle:
    ragged tile tails (a tile is 1024 elements), single blocks, zero blocks, fp32 and bf16 outputs, launch counters."""
    from llm_mixed_q_b200 import _lib as L
    from llm_mixed_q_b200.models.quantize.quantizers import minifloat_denorm_quantizer
    from llm_mixed_q_b200.models.quantize.quantizers.utils import canonicalise, launch_quantize, make_format

    lib = L.load()
    g = torch.Generator(device="cuda").manual_seed(99)
    names = [lib.bq_kernel_name(i).decode() for i in range(lib.bq_kernel_count())]
    sid, rid = names.index("quant_stream_kernel"), names.index("quant_rows_kernel")
    try:
        for n_blocks in (1, 2, 63, 64, 65, 127, 1000, 8 * 148 * 64 * 3 + 17):
            x = torch.randn(n_blocks, 16, device="cuda", generator=g) * 4
            x.view(-1)[::11] = 0
            x[::5] = 0                                   # all-zero blocks (block_log: tensor-global min)
            for name, kw in FORMATS:
                lib.bq_set_stream_quantizer(1)
                s0 = lib.bq_launch_count(sid)
                y, yo = both(name, kw, x, [1, 16], True)
                assert lib.bq_launch_count(sid) == s0 + 1, "dense block-16 input must take the streaming kernel"
                lib.bq_set_stream_quantizer(0)
                r0 = lib.bq_launch_count(rid)
                y2 = product(name)(x, block_size=[1, 16], skip_first_dim=True, **kw)
                assert lib.bq_launch_count(rid) == r0 + 1
                assert n_bits_diff(y, yo) == 0 and n_bits_diff(y2, yo) == 0, (n_blocks, name)
                # bf16 output of the same call (exact for <= 8 significant bits; block_log values are powers of two)
                lib.bq_set_stream_quantizer(1)
                canon = canonicalise(x, [1, 16], True, blocked=True)
                fkw = dict(kw)
                if "exponent_bias" in fkw and fkw["exponent_bias"] is None:
                    fkw["exponent_bias"] = 2 ** (fkw["exponent_width"] - 1) - 1
                fmt = make_format(name, b0=canon.b0, b1=canon.b1, fold=canon.fold, **fkw)
                yb = launch_quantize(x, fmt, canon, out_dtype=torch.bfloat16)
                assert torch.equal(yb.float(), yo.to(torch.bfloat16).float()), (n_blocks, name, "bf16")
        # element-wise kind: sizes that are multiples of 4 but not of 16
        for n in (4, 20, 1028, 4096 + 12, 3 * 1024 * 1184 + 4):
            x = torch.randn(n, device="cuda", generator=g)
            lib.bq_set_stream_quantizer(1)
            y = minifloat_denorm_quantizer(x, 8, 4, 7)
            assert n_bits_diff(y, O.minifloat_denorm_quantize(x, 8, 4, 7)) == 0, n
    finally:
        lib.bq_set_stream_quantizer(1)
