"""integer (fixed-point) quantizer — reference quantizers/integer.py:25-95; used for Llama's RoPE tables."""
from torch import Tensor

from .utils import quantize_elementwise


def integer_quantizer(x, width: int, frac_width: int, is_signed: bool = True):
    """clamp(round(x * 2^frac), int_min, int_max) / 2^frac (reference integer.py:77-95)."""
    if isinstance(x, int):
        return x
    if not isinstance(x, Tensor):
        scale = 2**frac_width
        lo, hi = (-(2 ** (width - 1)), 2 ** (width - 1) - 1) if is_signed else (0, 2**width - 1)
        return min(max(round(x * scale), lo), hi) / scale
    if not is_signed:
        raise NotImplementedError("unsigned integer quantisation is not on the accelerated path")
    return quantize_elementwise(x, "integer", width=width, exponent_bias=frac_width)
