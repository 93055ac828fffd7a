"""llm_mixed_q_b200.models — quantize/ (registries, kernels) and the quantized OPT / Llama / BERT module classes.

Same lookup surface as reference models/__init__.py:26-108 for the entries that sit on the quantized forward path
(model / config / quant-config-parser / analytic-profiler maps).  The tokenizer, sampler and statistic-hook maps belong to the
search / training drivers, which SURVEY.md §8 marks out of scope.
"""
from .bert_quantized import (BertQuantizedConfig, BertQuantizedForSequenceClassification, parse_bert_quantized_config,
                             profile_bert_quantized)
from .llama_quantized import (LlamaQuantizedConfig, LlamaQuantizedForCausalLM, LlamaQuantizedForSequenceClassification,
                              parse_llama_quantized_config, profile_llama_quantized)
from .opt_quantized import (OPTQuantizedConfig, OPTQuantizedForCausalLM, OPTQuantizedForSequenceClassification,
                            parse_opt_quantized_config, profile_opt_quantized)

MODEL_MAP = {
    "bert": {"cls": BertQuantizedForSequenceClassification},
    "llama": {"cls": LlamaQuantizedForSequenceClassification, "lm": LlamaQuantizedForCausalLM},
    "opt": {"cls": OPTQuantizedForSequenceClassification, "lm": OPTQuantizedForCausalLM},
}
CONFIG_MAP = {"bert": BertQuantizedConfig, "llama": LlamaQuantizedConfig, "opt": OPTQuantizedConfig}
QUANT_CONFIG_PARSER_MAP = {"bert": parse_bert_quantized_config, "llama": parse_llama_quantized_config,
                           "opt": parse_opt_quantized_config}
PROFILER_MAP = {"bert": profile_bert_quantized, "llama": profile_llama_quantized, "opt": profile_opt_quantized}


def get_model_cls(arch: str, task: str):
    assert arch in MODEL_MAP, f"arch {arch} not supported"
    assert task in MODEL_MAP[arch], f"task {task} not supported for arch {arch}"
    return MODEL_MAP[arch][task]


def get_config_cls(arch: str):
    assert arch in CONFIG_MAP, f"arch {arch} not supported"
    return CONFIG_MAP[arch]


def get_quant_config_parser(arch: str):
    assert arch in QUANT_CONFIG_PARSER_MAP, f"arch {arch} not supported"
    return QUANT_CONFIG_PARSER_MAP[arch]


def get_model_profiler(arch: str):
    assert arch in PROFILER_MAP, f"arch {arch} not supported"
    return PROFILER_MAP[arch]
