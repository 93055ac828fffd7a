from .configuration_opt import OPTQuantizedConfig
from .modeling_opt import (OPTQuantizedDecoder, OPTQuantizedDecoderLayer, OPTQuantizedForCausalLM,
                           OPTQuantizedForSequenceClassification, OPTQuantizedModel, OPTQauntizedAttention)
from .quant_config_opt import parse_opt_quantized_config
from .profiler_opt import profile_opt_quantized
