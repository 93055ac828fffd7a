"""Analytic profile of the quantized Llama decoder stack — mirror of reference models/llama_quantized/profiler_llama.py:9-153
(`profile_llama_quantized(config, seq_len)`): per layer q/k/v_proj, matmul_0 and matmul_1 once per head, o_proj, gate / down /
up_proj; no biases."""
from ..quantize.quantized_layer_profiler import profile_transformer_layers


def profile_llama_quantized(config, seq_len: int) -> dict:
    H, I, heads = config.hidden_size, config.intermediate_size, config.num_attention_heads
    d = H // heads

    def ops(lq):
        at, mlp = lq["self_attn"], lq["mlp"]
        for name in ("q_proj", "k_proj", "v_proj"):
            yield ("linear", at[name], H, H, False)
        for _ in range(heads):
            yield ("matmul", at["matmul_0"], (seq_len, d), (d, seq_len))
            yield ("matmul", at["matmul_1"], (seq_len, seq_len), (seq_len, d))
        yield ("linear", at["o_proj"], H, H, False)
        yield ("linear", mlp["gate_proj"], H, I, False)
        yield ("linear", mlp["down_proj"], I, H, False)
        yield ("linear", mlp["up_proj"], H, I, False)

    return profile_transformer_layers(config, seq_len, ops)
