"""LlamaQuantizedConfig — HF LlamaConfig (v4.31-era fields) + `quant_config`, parsed on assignment
(reference llama_quantized/configuration_llama.py:112-157)."""
from transformers.configuration_utils import PretrainedConfig

from .quant_config_llama import parse_llama_quantized_config


class LlamaQuantizedConfig(PretrainedConfig):
    model_type = "llama"
    keys_to_ignore_at_inference = ["past_key_values"]

    def __init__(
        self,
        vocab_size=32000,
        hidden_size=4096,
        intermediate_size=11008,
        num_hidden_layers=32,
        num_attention_heads=32,
        hidden_act="silu",
        max_position_embeddings=2048,
        initializer_range=0.02,
        rms_norm_eps=1e-6,
        use_cache=True,
        pad_token_id=0,
        bos_token_id=1,
        eos_token_id=2,
        tie_word_embeddings=False,
        quant_config=None,
        **kwargs,
    ):
        self.vocab_size = vocab_size
        self.max_position_embeddings = max_position_embeddings
        self.hidden_size = hidden_size
        self.intermediate_size = intermediate_size
        self.num_hidden_layers = num_hidden_layers
        self.num_attention_heads = num_attention_heads
        self.hidden_act = hidden_act
        self.initializer_range = initializer_range
        self.rms_norm_eps = rms_norm_eps
        self.use_cache = use_cache
        self.quant_config = quant_config
        self.pad_token_id = pad_token_id
        self.bos_token_id = bos_token_id
        self.eos_token_id = eos_token_id
        self.tie_word_embeddings = tie_word_embeddings
        super().__init__(**kwargs)

    def __setattr__(self, key, value):
        if key == "quant_config" and value is not None:
            value = parse_llama_quantized_config(config=value, num_hidden_layers=self.num_hidden_layers)
        return super().__setattr__(key, value)
