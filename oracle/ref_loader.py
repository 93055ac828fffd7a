"""
ORACLE — TEST INFRASTRUCTURE ONLY.

Imports the UNMODIFIED reference (read-only, /root/reference/src) with the
parent-package stubbing recipe of SURVEY.md Appendix C, so that
oracle/gen_golden.py can run the reference itself in the authoring container and
write golden vectors to tests/golden/.  /root/reference does not exist on the GPU
box: nothing executed there may import this module (`available()` says so).
"""
import importlib
import importlib.machinery
import os
import sys
import types
import warnings

SRC = os.environ.get("BQ_REFERENCE_SRC", "/root/reference/src")


def available() -> bool:
    return os.path.isdir(os.path.join(SRC, "llm_mixed_q", "models", "quantize"))


def _stub(name, path):
    if name not in sys.modules:
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m


def load_quantize():
    """Returns the reference `llm_mixed_q.models.quantize` sub-modules (quantizers, modules, functions, parser)."""
    if not available():
        raise RuntimeError("reference tree not present (expected in the authoring container only)")
    warnings.filterwarnings("ignore", message="Using a non-tuple sequence")
    if SRC not in sys.path:
        sys.path.insert(0, SRC)
    base = f"{SRC}/llm_mixed_q"
    _stub("llm_mixed_q", base)
    _stub("llm_mixed_q.models", f"{base}/models")
    _stub("llm_mixed_q.models.quantize", f"{base}/models/quantize")
    ns = types.SimpleNamespace()
    ns.quantizers = importlib.import_module("llm_mixed_q.models.quantize.quantizers")
    ns.modules = importlib.import_module("llm_mixed_q.models.quantize.quantized_modules")
    ns.functions = importlib.import_module("llm_mixed_q.models.quantize.quantized_functions")
    ns.parser = importlib.import_module("llm_mixed_q.models.quantize.quant_config_parser")
    return ns


def load_models():
    """Reference OPT / Llama quantized model + config classes (needs transformers; optuna is shimmed)."""
    if not available():
        raise RuntimeError("reference tree not present")
    import transformers  # noqa: F401  (real one first)

    warnings.filterwarnings("ignore", message="Using a non-tuple sequence")
    if "optuna" not in sys.modules:
        opt = types.ModuleType("optuna")
        opt.Trial = object
        opt.__spec__ = importlib.machinery.ModuleSpec("optuna", None)
        sys.modules["optuna"] = opt
    if SRC not in sys.path:
        sys.path.insert(0, SRC)
    base = f"{SRC}/llm_mixed_q"
    _stub("llm_mixed_q", base)
    _stub("llm_mixed_q.models", f"{base}/models")
    _stub("llm_mixed_q.utils", f"{base}/utils")
    for fam in ("opt", "llama"):
        _stub(f"llm_mixed_q.models.{fam}_quantized", f"{base}/models/{fam}_quantized")
    q = sys.modules.get("llm_mixed_q.models.quantize")
    if q is not None and not hasattr(q, "get_quantized_cls"):
        del sys.modules["llm_mixed_q.models.quantize"]      # drop load_quantize()'s stub, run the real __init__
    importlib.import_module("llm_mixed_q.models.quantize")
    ns = types.SimpleNamespace()
    ns.opt_cfg = importlib.import_module("llm_mixed_q.models.opt_quantized.configuration_opt")
    ns.opt = importlib.import_module("llm_mixed_q.models.opt_quantized.modeling_opt")
    ns.opt_qc = importlib.import_module("llm_mixed_q.models.opt_quantized.quant_config_opt")
    ns.llama_cfg = importlib.import_module("llm_mixed_q.models.llama_quantized.configuration_llama")
    ns.llama = importlib.import_module("llm_mixed_q.models.llama_quantized.modeling_llama")
    ns.llama_qc = importlib.import_module("llm_mixed_q.models.llama_quantized.quant_config_llama")
    return ns


def load_bert():
    """Reference BERT quantized model + config (after load_models()).  transformers 5.5 removed two helpers the
    reference's 4.31-era file expects; both are shimmed WITHOUT touching the reference: `find_pruneable_heads_and_indices`
    (import-time only, head pruning is never called) and `get_head_mask` (returns [None]*L when no head mask is given)."""
    ns = load_models()
    import transformers.pytorch_utils as pu

    if not hasattr(pu, "find_pruneable_heads_and_indices"):
        def _unsupported(*a, **k):
            raise NotImplementedError("head pruning is not part of the golden run")
        pu.find_pruneable_heads_and_indices = _unsupported
    _stub("llm_mixed_q.models.bert_quantized", f"{SRC}/llm_mixed_q/models/bert_quantized")
    ns.bert_cfg = importlib.import_module("llm_mixed_q.models.bert_quantized.configuration_bert")
    ns.bert = importlib.import_module("llm_mixed_q.models.bert_quantized.modeling_bert")
    ns.bert_qc = importlib.import_module("llm_mixed_q.models.bert_quantized.quant_config_bert")
    if not hasattr(ns.bert.BertQuantizedModel, "get_head_mask"):
        ns.bert.BertQuantizedModel.get_head_mask = lambda self, head_mask, n, *a, **k: [None] * n
    return ns
