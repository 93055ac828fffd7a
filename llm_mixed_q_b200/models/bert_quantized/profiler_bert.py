"""Analytic profile of the quantized BERT encoder stack — mirror of reference models/bert_quantized/profiler_bert.py:22-180
(`profile_bert_quantized(config, seq_len)`): per layer query / key / value, matmul_0 and matmul_1 once per head,
attention.output.dense, intermediate.dense, output.dense; all with bias (cross-attention is not profiled there either)."""
from ..quantize.quantized_layer_profiler import profile_transformer_layers


def profile_bert_quantized(config, seq_len: int) -> dict:
    H, I, heads = config.hidden_size, config.intermediate_size, config.num_attention_heads
    d = H // heads

    def ops(lq):
        at = lq["attention"]
        for name in ("query", "key", "value"):
            yield ("linear", at[name], H, H, True)
        for _ in range(heads):
            yield ("matmul", at["matmul_0"], (seq_len, d), (d, seq_len))
            yield ("matmul", at["matmul_1"], (seq_len, seq_len), (seq_len, d))
        yield ("linear", at["output"]["dense"], H, H, True)
        yield ("linear", lq["intermediate"]["dense"], H, I, True)
        yield ("linear", lq["output"]["dense"], I, H, True)

    return profile_transformer_layers(config, seq_len, ops)
