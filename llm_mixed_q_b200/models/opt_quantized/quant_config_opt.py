"""OPT quant-config expansion — same result as reference opt_quantized/quant_config_opt.py:34-97."""
from ..quant_config_expand import parse_model_quant_config

OPT_LAYER_TEMPLATE = {
    "self_attn": {
        "q_proj": "linear",
        "k_proj": "linear",
        "v_proj": "linear",
        "out_proj": "linear",
        "bmm_0": "matmul",
        "bmm_1": "matmul",
    },
    "fc1": "linear",
    "fc2": "linear",
}


def parse_opt_quantized_config(config, num_hidden_layers: int, strict: bool = True) -> dict:
    return parse_model_quant_config(config, num_hidden_layers, OPT_LAYER_TEMPLATE, {"linear": "linear", "matmul": "bmm"},
                                    strict=strict)
