"""BERT quant-config expansion — same result as reference bert_quantized/quant_config_bert.py:56-130
(`[linear]` / `[matmul]` type sections, `[model_layer]`, `[model_layer_<i>.attention.query]` …)."""
from ..quant_config_expand import parse_model_quant_config

BERT_LAYER_TEMPLATE = {
    "attention": {
        "query": "linear",
        "key": "linear",
        "value": "linear",
        "matmul_0": "matmul",
        "matmul_1": "matmul",
        "output": {"dense": "linear"},
    },
    "intermediate": {"dense": "linear"},
    "output": {"dense": "linear"},
}


def parse_bert_quantized_config(config, num_hidden_layers: int, strict: bool = True) -> dict:
    return parse_model_quant_config(config, num_hidden_layers, BERT_LAYER_TEMPLATE,
                                    {"linear": "linear", "matmul": "matmul"}, strict=strict)
