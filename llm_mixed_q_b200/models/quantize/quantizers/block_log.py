"""block_log (BL) quantizer — reference quantizers/block_log.py:23-120 + log.py:22-56, SURVEY.md App. A.4."""
from torch import Tensor

from .utils import quantize_blocked


def block_log_quantizer(
    x: Tensor,
    width: int,
    exponent_bias_width: int = None,
    block_size=[16],
    skip_first_dim: bool = False,
):
    """
    Per-block shared bias from ceil(log2 max), power-of-two value per element; cannot represent 0 and the
    value all-zero blocks map to depends on the tensor-wide smallest non-zero block max (block_log.py:50-53).
    """
    if exponent_bias_width is None:
        # reference: 2**None raises TypeError at block_log.py:57
        raise TypeError("unsupported operand type(s) for ** or pow(): 'int' and 'NoneType'")
    return quantize_blocked(
        x, "block_log", block_size, skip_first_dim, width=width, exponent_bias_width=exponent_bias_width
    )
