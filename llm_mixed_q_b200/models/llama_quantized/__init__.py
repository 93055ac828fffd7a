from .configuration_llama import LlamaQuantizedConfig
from .modeling_llama import (LlamaQuantizedAttention, LlamaQuantizedDecoderLayer, LlamaQuantizedForCausalLM,
                             LlamaQuantizedForSequenceClassification, LlamaQuantizedMLP, LlamaQuantizedModel)
from .quant_config_llama import parse_llama_quantized_config
from .profiler_llama import profile_llama_quantized
