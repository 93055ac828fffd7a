"""Llama quant-config expansion — same result as reference llama_quantized/quant_config_llama.py:38-116."""
from ..quant_config_expand import parse_model_quant_config

LLAMA_LAYER_TEMPLATE = {
    "self_attn": {
        "q_proj": "linear",
        "k_proj": "linear",
        "v_proj": "linear",
        "o_proj": "linear",
        "rotary_positional_encoding": "rotary_positional_encoding",
        "matmul_0": "matmul",
        "matmul_1": "matmul",
    },
    "mlp": {
        "gate_proj": "linear",
        "down_proj": "linear",
        "up_proj": "linear",
    },
}


def parse_llama_quantized_config(config, num_hidden_layers: int, strict: bool = True) -> dict:
    return parse_model_quant_config(
        config, num_hidden_layers, LLAMA_LAYER_TEMPLATE,
        {"linear": "linear", "rotary_positional_encoding": "rotary_positional_encoding", "matmul": "matmul"},
        strict=strict)
