"""block_minifloat (BM) quantizer — reference quantizers/block_minifloat.py:22-141, SURVEY.md App. A.3."""
from torch import Tensor

from .utils import quantize_blocked


def block_minifloat_quantizer(
    x: Tensor,
    width: int,
    exponent_width: int,
    exponent_bias_width: int,
    block_size=[16],
    skip_first_dim: bool = False,
):
    """
    Per-block shared exponent bias clamp(floor(log2 max), 0, 2^bw-1), IEEE-style minifloat per element
    (reference `block_minifloat_quantizer`, block_minifloat.py:110-141, inner minifloat.py:134-196).
    """
    return quantize_blocked(
        x, "block_minifloat", block_size, skip_first_dim,
        width=width, exponent_width=exponent_width, exponent_bias_width=exponent_bias_width,
    )
