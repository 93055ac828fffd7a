"""
ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product package.

A restatement of llm-mixed-q's software-emulated block quantisation hot path
(reference @ 740bf48, paths relative to /root/reference/src/llm_mixed_q/models/quantize/).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this file.

Why torch and not numpy: the reference's arithmetic *is* torch's elementwise fp32
arithmetic (torch.log2 / torch.round / 2**t / isclose).  numpy's float32 log2 is a
different libm and disagrees with torch in the last ulp at the ceil/floor/round
cliffs, so a numpy port could not be pinned bit-exactly.
What is restated here is the *algorithm*: the reference's ~45-op chain with
F.unfold/F.fold blocking is replaced by a pad + reshape blocking and one fused
expression per format.  The restatement is device agnostic: on CPU it is pinned
against the imported reference (tests/golden/, oracle/gen_golden.py); on a B200
the same expressions run on torch-CUDA and are what the reference itself would
compute after `.to("cuda")`.

Parity status: PINNED — bit-exact against the imported reference on every case in
tests/golden/manifest.json (generated here by oracle/gen_golden.py); the reference
itself ships no tests or golden vectors, so the pin is "reference run
in the authoring container", not "reference's own KATs".
"""
from __future__ import annotations

from math import ceil

import torch
import torch.nn.functional as F
from torch import Tensor

# ---------------------------------------------------------------------------
# blocking  (quantizers/utils.py:42-321)
# ---------------------------------------------------------------------------


def infer_block_shape(x_shape, block_shape):
    """quantizers/utils.py:42-67 — right-align, -1 / oversize ⇒ whole dim."""
    x_shape = list(x_shape)
    block_shape = list(block_shape)
    nd = len(x_shape)
    if len(block_shape) >= nd:
        out = block_shape[-nd:] if nd > 0 else []
    else:
        out = [-1] * (nd - len(block_shape)) + block_shape
    out = list(out)
    for i in range(nd):
        if out[i] == -1 or out[i] > x_shape[i]:
            out[i] = x_shape[i]
    return out


def _canon3(x: Tensor, block_size, skip_first_dim: bool):
    """
    Canonicalise every blocking case of utils.py:261-284 to a 3-D problem
    [L, R, C] with a (b0, b1) block over the last two dims and L never blocked.
    The third return value says whether the reference un-blocks this case with
    F.fold (utils.py:186-208, :240-258): col2im accumulates into a zero-filled
    buffer, so a -0.0 coming out of the element math becomes +0.0 there, whereas
    the reshape-based 1-D / 2-D-activation paths keep -0.0.

      1-D bias        [N]      -> [1, 1, N],  block (1, b)          utils.py:86-104
      2-D activation  [B, H]   -> [B, 1, H],  block (1, b1)         utils.py:127-144
      2-D weight      [R, C]   -> [1, R, C],  block (b0, b1)        utils.py:161-183
      3-D activation  [B,R,C]  -> [B, R, C],  block (b0, b1)        utils.py:211-237
    """
    if isinstance(block_size, int):
        block_size = [block_size]
    block_size = list(block_size)
    if x.ndim == 1:
        assert skip_first_dim is False, "skip_first_dim must be False for bias to be blocked"
        (b,) = infer_block_shape(list(x.shape), block_size)
        return x.reshape(1, 1, -1), (1, b), False
    if x.ndim == 2:
        if skip_first_dim:
            bs = infer_block_shape([1, x.shape[1]], block_size)
            return x.reshape(x.shape[0], 1, x.shape[1]), (1, bs[1]), False
        bs = infer_block_shape(list(x.shape), block_size)
        return x.reshape(1, x.shape[0], x.shape[1]), (bs[0], bs[1]), True
    if x.ndim == 3:
        if not skip_first_dim:
            raise NotImplementedError("block 3d weight is not supported.")
        bs = infer_block_shape([1, x.shape[1], x.shape[2]], block_size)
        return x, (bs[1], bs[2]), True
    raise RuntimeError(f"Unsupported x.ndim = {x.ndim}")


def block_view(x: Tensor, block_size, skip_first_dim: bool):
    """
    Returns (xb, meta): xb is [L, nb0, nb1, b0*b1] (zero padded on the right of
    each blocked dim, utils.py:70-83) and meta lets `unblock_view` undo it.
    """
    x3, (b0, b1), folds = _canon3(x, block_size, skip_first_dim)
    L, R, C = x3.shape
    b0 = max(int(b0), 1)
    b1 = max(int(b1), 1)
    Rp = ceil(R / b0) * b0 if R else 0
    Cp = ceil(C / b1) * b1 if C else 0
    xp = F.pad(x3, (0, Cp - C, 0, Rp - R))
    nb0, nb1 = (Rp // b0 if b0 else 0), (Cp // b1 if b1 else 0)
    xb = xp.reshape(L, nb0, b0, nb1, b1).permute(0, 1, 3, 2, 4).reshape(L, nb0, nb1, b0 * b1)
    return xb, (tuple(x.shape), L, R, C, b0, b1, nb0, nb1, folds)


def unblock_view(yb: Tensor, meta):
    shape, L, R, C, b0, b1, nb0, nb1, folds = meta
    y = yb.reshape(L, nb0, nb1, b0, b1).permute(0, 1, 3, 2, 4).reshape(L, nb0 * b0, nb1 * b1)
    if folds:
        y = y + 0.0          # col2im: 0 + (-0.0) = +0.0
    return y[:, :R, :C].reshape(shape)


def fixed_block_max(xb: Tensor):
    """
    per-block max |x| with the zero substitution every block quantizer applies
    (block_fp.py:53-58, block_minifloat.py:51-55, block_log.py:49-53): if every
    block max is 0 → ones, else zero maxima := the smallest non-zero block max of
    the whole tensor call.
    """
    m = xb.abs().amax(dim=-1, keepdim=True)
    if m.numel() == 0:
        return m
    if torch.all(m == 0):
        return torch.ones_like(m)
    m = m.clone()
    m[m == 0] = m[m != 0].min()
    return m


def _default_bias(exponent_bias, exponent_width):
    if exponent_bias in (None, "none", "None"):
        return 2 ** (exponent_width - 1) - 1
    return exponent_bias


def _zero_like_ref(x):
    return torch.tensor([0.0], dtype=x.dtype, device=x.device)


# ---------------------------------------------------------------------------
# element formats
# ---------------------------------------------------------------------------


def block_fp_quantize(x, width=12, exponent_width=8, exponent_bias=None, block_size=(16,), skip_first_dim=True):
    """block_fp.py:21-96."""
    xb, meta = block_view(x, block_size, skip_first_dim)
    mx = fixed_block_max(xb)
    m = width - 1
    bias = _default_bias(exponent_bias, exponent_width)
    emax, emin = 2**exponent_width - 1 - bias, -bias
    sign = torch.sign(xb + 1e-9)                                    # block_fp.py:69
    value = torch.abs(xb) + 1e-9                                    # :71
    e = torch.ceil(torch.log2(mx)).clamp(min=emin, max=emax)       # :72-73
    mant = value / 2**e                                             # :75
    shift = 2**m
    q = torch.round(mant * shift).clamp(min=0, max=2**m - 1)       # :77-79
    mant = q / shift                                                # :80
    yb = sign * (2**e) * mant                                       # :82
    y = unblock_view(yb, meta)
    c = torch.isclose(x, _zero_like_ref(x))                         # :93
    return (~c) * y + c * x                                         # :94


def minifloat_ieee_quantize(x, width, exponent_width, exponent_bias=None):
    """minifloat.py:134-196; `exponent_bias` may be a tensor (block_minifloat.py:60-65)."""
    M = width - exponent_width - 1
    exponent_bias = _default_bias(exponent_bias, exponent_width)
    emax = 2**exponent_width - 1 - exponent_bias
    emin = -exponent_bias
    shift = 2**M
    sign = torch.sign(x + 1e-9)                                     # :172
    value = torch.abs(x)
    e = torch.floor(torch.log2(value + 1e-9))                       # :176
    e = e.clamp(min=emin, max=emax)                                 # my_clamp :177 (tensor or scalar bounds)
    mant = value / 2**e                                             # :179
    if isinstance(exponent_bias, (int, float)):
        exponent_bias = torch.tensor([exponent_bias], dtype=e.dtype, device=e.device)
    normal = ~torch.isclose(e, -exponent_bias)                      # :187
    sm = normal * torch.round(mant * shift - shift).clamp(min=0, max=2**M - 1) + (~normal) * torch.round(
        mant * shift / 2
    ).clamp(min=0, max=2**M - 1)                                    # :189-190
    mant = normal * (1.0 + sm / shift) + (~normal) * (sm / shift * 2)  # :191
    c = torch.isclose(value, _zero_like_ref(value))                 # :193
    return (~c) * (sign * (2**e) * mant) + c * x                    # :194


def block_minifloat_quantize(x, width, exponent_width, exponent_bias_width, block_size=(16,), skip_first_dim=False):
    """block_minifloat.py:22-74."""
    xb, meta = block_view(x, block_size, skip_first_dim)
    mx = fixed_block_max(xb)
    b = torch.floor(torch.log2(mx)).clamp(min=0, max=2**exponent_bias_width - 1)  # :57-59
    yb = minifloat_ieee_quantize(xb, width, exponent_width, b)
    return unblock_view(yb, meta)


def log_quantize(x, width, exponent_bias):
    """log.py:22-56; `exponent_bias` may be a tensor (block_log.py:60)."""
    eb = width - 1
    if not isinstance(exponent_bias, Tensor) and exponent_bias in (None, "none", "None"):
        exponent_bias = 2 ** (eb - 1) - 1
    emax = 2**eb - 1 - exponent_bias
    emin = -exponent_bias
    min_pos = 2**emin                                               # :49
    sign = torch.sign(x + min_pos * 0.1)                            # :51
    value = torch.abs(x) + min_pos * 0.1                            # :52
    e = torch.round(torch.log2(value))                              # :54
    e = e.clamp(min=emin, max=emax)
    return sign * (2**e)                                            # :56


def block_log_quantize(x, width, exponent_bias_width=None, block_size=(16,), skip_first_dim=False):
    """block_log.py:23-69."""
    eb = width - 1
    xb, meta = block_view(x, block_size, skip_first_dim)
    mx = fixed_block_max(xb)
    me = torch.ceil(torch.log2(mx))                                 # :55
    b = (2**eb - 1 - me).clamp(min=0, max=2**exponent_bias_width - 1)  # :56-58
    yb = log_quantize(xb, width, b)
    return unblock_view(yb, meta)


def minifloat_denorm_quantize(x, width, exponent_width, exponent_bias=None):
    """minifloat.py:21-82."""
    M = width - exponent_width - 1
    bias = _default_bias(exponent_bias, exponent_width)
    emax, emin = 2**exponent_width - 1 - bias, -bias
    sign = torch.sign(x + 1e-9)                                     # :60
    value = torch.abs(x)
    e = torch.ceil(torch.log2(value + 1e-9)).clamp(min=emin, max=emax)  # :64-65
    mant = value / 2**e                                             # :69
    shift = 2**M
    q = torch.round(mant * shift).clamp(min=0, max=2**M - 1)       # :71-75
    mant = q / shift
    c = torch.isclose(value, _zero_like_ref(value))                 # :79
    return (~c) * (sign * (2**e) * mant) + c * x                    # :80


def integer_quantize(x, width, frac_width, is_signed=True):
    """integer.py:25-58 (adjacent row f3: Llama RoPE tables)."""
    if is_signed:
        lo, hi = -(2 ** (width - 1)), 2 ** (width - 1) - 1
    else:
        lo, hi = 0, 2**width - 1
    scale = 2**frac_width
    return torch.round(x.mul(scale)).clamp(min=lo, max=hi).div(scale)


QUANTIZERS = {
    "block_fp": block_fp_quantize,
    "block_minifloat": block_minifloat_quantize,
    "block_log": block_log_quantize,
    "minifloat_denorm": minifloat_denorm_quantize,
    "minifloat_ieee": minifloat_ieee_quantize,
    "integer": integer_quantize,
}

# ---------------------------------------------------------------------------
# consumers: Linear (quantized_modules/linear.py:59-76) and matmul/bmm
# (quantized_functions/matmul.py:146-297)
# ---------------------------------------------------------------------------

_FMT_KEYS = {
    "block_fp": ("width", "exponent_width", "exponent_bias", "block_size"),
    "block_minifloat": ("width", "exponent_width", "exponent_bias_width", "block_size"),
    "block_log": ("width", "exponent_bias_width", "block_size"),
    "minifloat_denorm": ("width", "exponent_width", "exponent_bias"),
}


def operand_quantizer(config: dict, prefix: str, skip_first_dim: bool):
    """Bind a quantizer from `<prefix>_*` config keys the way linear.py:113-281 / matmul.py do."""
    name = config["name"]
    if name == "log":          # quantized_functions/__init__.py:20,29 aliases log -> block_log
        name = "block_log"
    kw = {k: config[f"{prefix}_{k}"] for k in _FMT_KEYS[name]}
    fn = QUANTIZERS[name]
    if "block_size" in kw:
        return lambda t: fn(t, skip_first_dim=skip_first_dim, **kw)
    return lambda t: fn(t, **kw)


# ---------------------------------------------------------------------------
# noise-floor control: the same restatement with the contraction run in the opposite order.  Mathematically identical;
# on a real BLAS the fp32 partial sums round differently, which is the ONLY perturbation — what tests use to measure how far
# two correct implementations of the quantised forward drift apart (rounding flips seeded by accumulation order).
# ---------------------------------------------------------------------------
ACCUMULATION_ORDER = "natural"          # or "reversed"


def gemm_linear(x, w, b=None):
    if ACCUMULATION_ORDER == "reversed":
        return F.linear(x.flip(-1).contiguous(), w.flip(-1).contiguous(), b)
    return F.linear(x, w, b)


def gemm_mm(mm, x, y):
    if ACCUMULATION_ORDER == "reversed":
        return mm(x.flip(-1).contiguous(), y.flip(-2).contiguous())
    return mm(x, y)


def linear_forward(x, weight, bias, config):
    """
    _LinearBase.forward (linear.py:59-76), functional form: returns
    (y, weight_q, bias_q).  In PTQ mode the reference overwrites the parameters
    with weight_q / bias_q on the first call; steady state is
    F.linear(Qx(x), weight_q, bias_q).
    """
    if config.get("bypass", False):
        return F.linear(x, weight, bias), weight, bias
    xq = operand_quantizer(config, "data_in", True)(x)
    wq = operand_quantizer(config, "weight", False)(weight)
    bq = operand_quantizer(config, "bias", False)(bias) if bias is not None else None
    return gemm_linear(xq, wq, bq), wq, bq


def matmul_forward(x, y, config, style="matmul"):
    """generic_matmul_* (matmul.py:146-297). block_log quantizes x only (:293-296)."""
    mm = {"matmul": torch.matmul, "bmm": torch.bmm}[style]
    if config.get("bypass", False):
        return mm(x, y)
    name = "block_log" if config["name"] == "log" else config["name"]
    blocked = name in ("block_fp", "block_minifloat", "block_log")
    xs, ys = list(x.shape), list(y.shape)
    if blocked:
        xm, ym = x.ndim > 2, y.ndim > 2
        if xm:
            x = torch.flatten(x, 0, -3)
        if ym:
            y = torch.flatten(y, 0, -3)
        x = operand_quantizer(config, "data_in", xm)(x)
        if name != "block_log":
            y = operand_quantizer(config, "weight", ym)(y)
        x, y = x.reshape(xs), y.reshape(ys)
    else:
        x = operand_quantizer(config, "data_in", False)(x)
        y = operand_quantizer(config, "weight", False)(y)
    return gemm_mm(mm, x, y)


def perplexity_from_losses(losses, batch_size, seq_len):
    """eval/eval_lm.py:41-63 — ppl = exp( Σ loss·B·S / (S·N) ), N = B·num_batches."""
    import math

    tot = sum(float(l) * batch_size * seq_len for l in losses)
    n = batch_size * len(losses)
    return math.exp(tot / (seq_len * n))
