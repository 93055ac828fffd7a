"""BERT module classes + the OPT/Llama/BERT sequence-classification heads.

CPU: BERT quant-config expansion reproduces the reference's parser output for every shipped TOML and a per-layer mixed
config (tests/golden/configs_bert.json, written by oracle/gen_golden_bert.py from the unmodified reference); our classes
expose the reference's parameter names; the models registry answers like reference models/__init__.py.
GPU: tiny random-init models with the reference's weights vs the reference's own CPU forward.  Rounding is discontinuous,
so an ulp-level GEMM accumulation-order difference can flip an element by one quantisation step downstream — hidden states
are compared statistically (bounds in the asserts), the same way as tests/test_gpu_models.py."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLD

BERT_KW = dict(vocab_size=512, hidden_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=128,
               max_position_embeddings=64, num_labels=3)


def clone(d):
    return json.loads(json.dumps(d))


@pytest.fixture(scope="module")
def bert_cfgs():
    with open(os.path.join(GOLD, "configs_bert.json")) as f:
        return json.load(f)


def test_bert_expansion_matches_reference(bert_cfgs, golden_configs):
    from llm_mixed_q_b200.models.bert_quantized import parse_bert_quantized_config

    n = 0
    for fn, raw in golden_configs["raw"].items():
        exp = bert_cfgs["bert"][fn]
        if "error" in exp:
            with pytest.raises(Exception) as ei:
                parse_bert_quantized_config(clone(raw), 2)
            assert type(ei.value).__name__ == exp["error"]
        else:
            assert parse_bert_quantized_config(clone(raw), 2) == exp, fn
            n += 1
    assert n >= 5
    got = parse_bert_quantized_config(clone(bert_cfgs["mixed_raw"]), 3)
    assert got == bert_cfgs["mixed_bert"]
    assert got["model_layer_2"]["attention"]["output"]["dense"]["data_in_width"] == 6     # layer 2 not listed -> default
    assert parse_bert_quantized_config(None, 2) is None


def test_registry_matches_reference_surface():
    import llm_mixed_q_b200.models as M

    assert M.get_model_cls("bert", "cls").__name__ == "BertQuantizedForSequenceClassification"
    assert M.get_model_cls("opt", "lm").__name__ == "OPTQuantizedForCausalLM"
    assert M.get_model_cls("llama", "cls").__name__ == "LlamaQuantizedForSequenceClassification"
    assert M.get_config_cls("bert").__name__ == "BertQuantizedConfig"
    assert M.get_quant_config_parser("llama").__name__ == "parse_llama_quantized_config"
    with pytest.raises(AssertionError):
        M.get_model_cls("bert", "lm")
    with pytest.raises(AssertionError):
        M.get_config_cls("gpt2")


def _build_bert(tag, golden_configs, bert_cfgs):
    from llm_mixed_q_b200.models.bert_quantized import BertQuantizedConfig, BertQuantizedForSequenceClassification

    qc = {"bert_tiny_bfp6": lambda: golden_configs["raw"]["bfp_6bit.toml"], "bert_tiny_mixed": lambda: bert_cfgs["mixed_raw"],
          "bert_tiny_bmf8": lambda: golden_configs["raw"]["block_minifloat.toml"],
          "bert_tiny_bl8": lambda: golden_configs["raw"]["block_log.toml"]}[tag]()
    z = np.load(os.path.join(GOLD, tag + ".npz"))
    cfg = BertQuantizedConfig(quant_config=clone(qc), initializer_range=float(z["init"]), **BERT_KW)
    model = BertQuantizedForSequenceClassification(cfg).eval()
    sd = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd::")}
    return model, sd, z


def test_bert_parameter_names_and_classes_match_reference(golden_configs, bert_cfgs):
    model, sd, _ = _build_bert("bert_tiny_mixed", golden_configs, bert_cfgs)
    assert set(model.state_dict().keys()) == set(sd.keys())
    missing, unexpected = model.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    layer = model.bert.encoder.layer[0]
    assert type(layer.attention.self.query).__name__ == "LinearBlockFP"
    assert layer.attention.self.query.config["data_in_width"] == bert_cfgs["mixed_bert"]["model_layer_0"]["attention"]["query"]["data_in_width"]
    assert layer.attention.self.quant_config["matmul_0"]["name"] == "block_fp"
    # quant_config is parsed on assignment (reference configuration_bert.py:183)
    assert "model_layer_1" in model.config.quant_config and "default" in model.config.quant_config


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["bert_tiny_bfp6", "bert_tiny_mixed", "bert_tiny_bmf8", "bert_tiny_bl8"])
def test_bert_tiny_matches_reference_forward(tag, golden_configs, bert_cfgs):
    model, sd, z = _build_bert(tag, golden_configs, bert_cfgs)
    model.load_state_dict(sd, strict=True)
    model = model.cuda()
    ids, am, tt = (torch.from_numpy(z[k]).cuda() for k in ("input_ids", "attention_mask", "token_type_ids"))
    with torch.no_grad():
        out = model(input_ids=ids, attention_mask=am, token_type_ids=tt, output_hidden_states=True)
    ref_h, ref_logits = torch.from_numpy(z["last_hidden"]), torch.from_numpy(z["logits"])
    h = out.hidden_states[-1].cpu()
    err = (h - ref_h).abs()
    spread = float(ref_h.std())
    assert float(err.mean()) <= 0.03 * spread and float(err.max()) <= 0.75 * spread, (float(err.mean()), float(err.max()), spread)
    lerr = (out.logits.cpu() - ref_logits).abs()
    scale = max(float(ref_logits.abs().max()), 1e-3)
    assert float(lerr.max()) <= 0.1 * scale, (float(lerr.max()), scale)


@pytest.mark.gpu
def test_bert_classification_loss_paths(golden_configs, bert_cfgs):
    model, sd, z = _build_bert("bert_tiny_bfp6", golden_configs, bert_cfgs)
    model.load_state_dict(sd, strict=True)
    model = model.cuda()
    ids, am = torch.from_numpy(z["input_ids"]).cuda(), torch.from_numpy(z["attention_mask"]).cuda()
    labels = torch.tensor([0, 2], device="cuda")
    with torch.no_grad():
        out = model(input_ids=ids, attention_mask=am, labels=labels)
    ref = torch.nn.functional.cross_entropy(out.logits, labels)
    assert model.config.problem_type == "single_label_classification"
    torch.testing.assert_close(out.loss, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("family", ["opt", "llama"])
def test_sequence_classification_matches_reference(family, golden_configs):
    z = np.load(os.path.join(GOLD, f"{family}_tiny_cls.npz"))
    qc = clone(golden_configs["raw"]["bfp_6bit.toml"])
    if family == "opt":
        from llm_mixed_q_b200.models.opt_quantized import OPTQuantizedConfig, OPTQuantizedForSequenceClassification

        cfg = OPTQuantizedConfig(hidden_size=64, num_hidden_layers=2, ffn_dim=128, num_attention_heads=4, vocab_size=512,
                                 max_position_embeddings=64, quant_config=qc, num_labels=2, pad_token_id=1)
        model = OPTQuantizedForSequenceClassification(cfg).eval()
    else:
        from llm_mixed_q_b200.models.llama_quantized import LlamaQuantizedConfig, LlamaQuantizedForSequenceClassification

        cfg = LlamaQuantizedConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                                   vocab_size=512, max_position_embeddings=64, quant_config=qc, num_labels=2, pad_token_id=0)
        model = LlamaQuantizedForSequenceClassification(cfg).eval()
    sd = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd::")}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not missing, missing
    model = model.cuda()
    ids, am, labels = (torch.from_numpy(z[k]).cuda() for k in ("input_ids", "attention_mask", "labels"))
    with torch.no_grad():
        out = model(input_ids=ids, attention_mask=am, labels=labels)
    ref_logits, ref_loss = torch.from_numpy(z["logits"]), float(z["loss"])
    assert out.logits.shape == ref_logits.shape
    assert float((out.logits.cpu() - ref_logits).abs().max()) <= 0.05 * max(float(ref_logits.abs().max()), 1e-3)
    assert abs(float(out.loss) - ref_loss) <= 5e-3 * abs(ref_loss)
