"""
ORACLE — TEST INFRASTRUCTURE ONLY.  Goldens for the analytic layer profiler: runs the UNMODIFIED reference functions
(models/quantize/quantized_layer_profiler.py, models/{opt,llama,bert}_quantized/profiler_*.py) on every shipped quantization
TOML they accept and on the mixed-precision config -> tests/golden/profiler.json.
Run in the authoring container only (needs /root/reference):  python oracle/gen_golden_profiler.py
"""
import json
import os
import sys
from copy import deepcopy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def plain(d):
    return {k: int(v) for k, v in d.items()}


def main():
    import importlib

    models = ref_loader.load_bert()
    prof = importlib.import_module("llm_mixed_q.models.quantize.quantized_layer_profiler")
    p_opt = importlib.import_module("llm_mixed_q.models.opt_quantized.profiler_opt")
    p_llama = importlib.import_module("llm_mixed_q.models.llama_quantized.profiler_llama")
    p_bert = importlib.import_module("llm_mixed_q.models.bert_quantized.profiler_bert")
    g = json.load(open(os.path.join(GOLD, "configs.json")))
    out = {"linear": [], "matmul": [], "models": []}
    for name, raw in sorted(g["raw"].items()):
        node = raw["default"]
        for (fin, fout, bias, bs) in [(768, 3072, True, 2048), (4096, 11008, False, 100), (50, 70, True, 3)]:
            try:
                r = plain(prof.profile_linear_layer(deepcopy(node), fin, fout, bias, bs))
            except Exception as e:                      # formats the reference's profiler does not know
                r = {"error": type(e).__name__}
            out["linear"].append({"toml": name, "args": [fin, fout, bias, bs], "result": r})
        for (s0, s1) in [((2048, 64), (64, 2048)), ((2048, 2048), (2048, 64)), ((100, 24), (24, 100))]:
            try:
                r = plain(prof.profile_matmul_layer(deepcopy(node), s0, s1))
            except Exception as e:
                r = {"error": type(e).__name__}
            out["matmul"].append({"toml": name, "args": [list(s0), list(s1)], "result": r})
    for tomlname in ("bfp_6bit.toml", "bfp_4bit.toml", "integer.toml", "bypass.toml", None):
        raw = deepcopy(g["raw"][tomlname]) if tomlname else deepcopy(g["mixed_raw"])
        for arch, cfg_cls, fn, kw in [
            ("opt", models.opt_cfg.OPTQuantizedConfig, p_opt.profile_opt_quantized,
             dict(hidden_size=256, num_hidden_layers=3, ffn_dim=1024, num_attention_heads=4)),
            ("llama", models.llama_cfg.LlamaQuantizedConfig, p_llama.profile_llama_quantized,
             dict(hidden_size=256, intermediate_size=688, num_hidden_layers=3, num_attention_heads=4)),
            ("bert", models.bert_cfg.BertQuantizedConfig, p_bert.profile_bert_quantized,
             dict(hidden_size=256, intermediate_size=1024, num_hidden_layers=3, num_attention_heads=4)),
        ]:
            if tomlname is None and arch != "opt":
                continue                                  # the mixed config carries OPT layer names
            try:
                cfg = cfg_cls(quant_config=deepcopy(raw), **kw)
                r = plain(fn(cfg, 128))
            except Exception as e:
                r = {"error": type(e).__name__}
            out["models"].append({"toml": tomlname, "arch": arch, "kw": kw, "seq_len": 128, "result": r})
    with open(os.path.join(GOLD, "profiler.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print(len(out["linear"]), len(out["matmul"]), len(out["models"]), "cases;",
          sum("error" in c["result"] for k in out for c in out[k]), "of them errors")


if __name__ == "__main__":
    main()
