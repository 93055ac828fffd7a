"""minifloat_denorm / minifloat_ieee quantizers — reference quantizers/minifloat.py:21-239, SURVEY.md App. A.5."""
from torch import Tensor

from .utils import default_bias, quantize_elementwise


def minifloat_denorm_quantizer(x: Tensor, width: int, exponent_width: int, exponent_bias: int = None):
    """Element-wise minifloat without the implicit leading one (reference minifloat.py:104-131)."""
    return quantize_elementwise(
        x, "minifloat_denorm", width=width, exponent_width=exponent_width,
        exponent_bias=default_bias(exponent_bias, exponent_width),
    )


def minifloat_ieee_quantizer(x: Tensor, width: int, exponent_width: int, exponent_bias: int = None):
    """Element-wise IEEE-style minifloat with subnormals (reference minifloat.py:199-239)."""
    return quantize_elementwise(
        x, "minifloat_ieee", width=width, exponent_width=exponent_width,
        exponent_bias=default_bias(exponent_bias, exponent_width),
    )
