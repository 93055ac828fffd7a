"""QUANTIZED_MODULE_MAP["linear"][name] — same registry shape as reference quantized_modules/__init__.py:5-15."""
from .linear import (LinearBlockFP, LinearBlockLog, LinearBlockMinifloat, LinearInteger, LinearMinifloatDenorm,
                     LinearMinifloatIEEE)

QUANTIZED_MODULE_MAP = {
    "linear": {
        "block_fp": LinearBlockFP,
        "integer": LinearInteger,
        "minifloat_ieee": LinearMinifloatIEEE,
        "minifloat_denorm": LinearMinifloatDenorm,
        "block_log": LinearBlockLog,
        "block_minifloat": LinearBlockMinifloat,
        # "log": the reference's LinearLog passes a kwarg its quantizer does not accept (linear.py:234-239 vs
        # log.py:72-76) and cannot run; not provided.
    },
}
