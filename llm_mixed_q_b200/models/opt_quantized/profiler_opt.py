"""Analytic profile of the quantized OPT decoder stack — mirror of reference models/opt_quantized/profiler_opt.py:11-148
(`profile_opt_quantized(config, seq_len)`): per layer q/k/v_proj, bmm_0 and bmm_1 once per head, out_proj, fc1, fc2."""
from ..quantize.quantized_layer_profiler import profile_transformer_layers


def profile_opt_quantized(config, seq_len: int) -> dict:
    H, F, heads, bias = config.hidden_size, config.ffn_dim, config.num_attention_heads, config.enable_bias
    d = H // heads

    def ops(lq):
        at = lq["self_attn"]
        for name in ("q_proj", "k_proj", "v_proj"):
            yield ("linear", at[name], H, H, bias)
        for _ in range(heads):
            yield ("matmul", at["bmm_0"], (seq_len, d), (d, seq_len))
            yield ("matmul", at["bmm_1"], (seq_len, seq_len), (seq_len, d))
        yield ("linear", at["out_proj"], H, H, bias)
        yield ("linear", lq["fc1"], H, F, bias)
        yield ("linear", lq["fc2"], F, H, bias)

    return profile_transformer_layers(config, seq_len, ops)
