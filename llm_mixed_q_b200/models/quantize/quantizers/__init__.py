"""QUANTIZER_MAP — name -> quantizer, same names as the reference registry (quantizers/__init__.py:8-16)."""
from .block_fp import block_fp_quantizer
from .block_log import block_log_quantizer
from .block_minifloat import block_minifloat_quantizer
from .integer import integer_quantizer
from .minifloat import minifloat_denorm_quantizer, minifloat_ieee_quantizer

QUANTIZER_MAP = {
    "block_fp": block_fp_quantizer,
    "block_log": block_log_quantizer,
    "block_minifloat": block_minifloat_quantizer,
    "integer": integer_quantizer,
    "minifloat_denorm": minifloat_denorm_quantizer,
    "minifloat_ieee": minifloat_ieee_quantizer,
}
# The reference also registers an element-wise "log" quantizer (log.py:72-88).  It is not on the hot path
# (the "log" *functions* alias block_log, quantized_functions/__init__.py:20,29) and is not provided here.
