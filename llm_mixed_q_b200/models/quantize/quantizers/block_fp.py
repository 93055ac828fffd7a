"""block_fp (BFP / MSFP) quantizer — reference quantizers/block_fp.py:21-153, SURVEY.md App. A.2."""
from torch import Tensor

from .utils import default_bias, quantize_blocked


def block_fp_quantizer(
    x: Tensor,
    width: int = 12,
    exponent_width: int = 8,
    exponent_bias: int = None,
    block_size=[16],
    skip_first_dim: bool = True,
):
    """
    Shared exponent per block, sign + (width-1)-bit mantissa per element; |x| <= 1e-8 passes through.
    Same signature, defaults and return contract (new fp32 tensor of x's shape, STE backward) as the
    reference's `block_fp_quantizer` (block_fp.py:127-153); bit-identical values, one sm_100a kernel.
    """
    return quantize_blocked(
        x, "block_fp", block_size, skip_first_dim,
        width=width, exponent_width=exponent_width, exponent_bias=default_bias(exponent_bias, exponent_width),
    )
