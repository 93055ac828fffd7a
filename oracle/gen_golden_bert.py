"""
ORACLE — TEST INFRASTRUCTURE ONLY.

Golden fixtures for the BERT module classes and the sequence-classification heads, produced by running the
UNMODIFIED reference (oracle/ref_loader.py) on CPU in the authoring container:

  tests/golden/configs_bert.json     reference-parsed BERT quant configs (every shipped TOML + a per-layer mixed one)
  tests/golden/bert_tiny_*.npz       tiny BertQuantizedForSequenceClassification: state dict (before PTQ), inputs with a
                                     padded sequence, last hidden state, pooled output, logits
  tests/golden/{opt,llama}_tiny_cls.npz   tiny OPT/Llama ForSequenceClassification: pooled logits and CE loss

Usage (authoring container only):  python oracle/gen_golden_bert.py
"""
import json
import os
import random
import sys
from copy import deepcopy

import numpy as np
import toml
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
REF_CFG_DIR = "/root/reference/experiments/emnlp/configs/quantization"
BERT_KW = dict(vocab_size=512, hidden_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=128,
               max_position_embeddings=64, num_labels=3)


def sd_arrays(model):
    return {"sd::" + k: v.detach().cpu().numpy().copy() for k, v in model.state_dict().items()}


def mixed_bert(base):
    rnd = random.Random(1)
    mixed = {"default": deepcopy(base)}
    for i in (0, 1):
        for path in ("attention.query", "attention.key", "attention.value", "attention.output.dense", "attention.matmul_0",
                     "attention.matmul_1", "intermediate.dense", "output.dense"):
            if i == 1 and path in ("intermediate.dense", "attention.matmul_1"):
                continue                              # unspecified -> type default
            node = deepcopy(base)
            node["data_in_width"] = rnd.choice([6, 5, 4])
            node["weight_width"] = rnd.choice([6, 5, 4, 3])
            node["bias_width"] = rnd.choice([6, 5, 4, 3])
            cur = mixed.setdefault(f"model_layer_{i}", {})
            parts = path.split(".")
            for p in parts[:-1]:
                cur = cur.setdefault(p, {})
            cur[parts[-1]] = node
    return mixed


def main():
    m = ref_loader.load_bert()
    cfgs = {fn: toml.load(os.path.join(REF_CFG_DIR, fn)) for fn in sorted(os.listdir(REF_CFG_DIR))}
    out = {"bert": {}}
    for fn, raw in cfgs.items():
        try:
            out["bert"][fn] = m.bert_qc.parse_bert_quantized_config(deepcopy(raw), 2)
        except Exception as e:
            out["bert"][fn] = {"error": type(e).__name__}
    mixed = mixed_bert(cfgs["bfp_6bit.toml"]["default"])
    out["mixed_raw"] = mixed
    out["mixed_bert"] = m.bert_qc.parse_bert_quantized_config(deepcopy(mixed), 3)
    with open(os.path.join(GOLD, "configs_bert.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)

    rs = np.random.RandomState(7)
    ids = torch.from_numpy(rs.randint(1, 512, size=(2, 48)).astype(np.int64))
    am = torch.ones(2, 48, dtype=torch.long)
    am[1, 40:] = 0
    ids[1, 40:] = 0
    tt = torch.zeros(2, 48, dtype=torch.long)
    tt[:, 24:] = 1
    # block_minifloat needs a scaled init (N(0,0.02) weights all quantise to 0)
    for tag, qc, init in [("bert_tiny_bfp6", cfgs["bfp_6bit.toml"], 0.02), ("bert_tiny_mixed", mixed, 0.02),
                          ("bert_tiny_bmf8", cfgs["block_minifloat.toml"], 1.5), ("bert_tiny_bl8", cfgs["block_log.toml"], 0.02)]:
        torch.manual_seed(0)
        cfg = m.bert_cfg.BertQuantizedConfig(quant_config=deepcopy(qc), initializer_range=init, is_decoder=False,
                                             add_cross_attention=False, chunk_size_feed_forward=0, **BERT_KW)
        model = m.bert.BertQuantizedForSequenceClassification(cfg).eval()
        arrs = sd_arrays(model)
        with torch.no_grad():
            o = model(input_ids=ids, attention_mask=am, token_type_ids=tt, output_hidden_states=True)
            pooled = model.bert.pooler(o.hidden_states[-1])
        arrs.update(input_ids=ids.numpy(), attention_mask=am.numpy(), token_type_ids=tt.numpy(), logits=o.logits.numpy().copy(),
                    last_hidden=o.hidden_states[-1].numpy().copy(), pooled=pooled.numpy().copy(), init=np.array(init))
        np.savez_compressed(os.path.join(GOLD, tag + ".npz"), **arrs)
        print(tag, o.logits.flatten().tolist())

    labels = torch.tensor([1, 0])
    torch.manual_seed(0)
    cfg = m.opt_cfg.OPTQuantizedConfig(hidden_size=64, num_hidden_layers=2, ffn_dim=128, num_attention_heads=4, vocab_size=512,
                                       max_position_embeddings=64, quant_config=deepcopy(cfgs["bfp_6bit.toml"]), num_labels=2,
                                       pad_token_id=1)
    model = m.opt.OPTQuantizedForSequenceClassification(cfg).eval()
    arrs = sd_arrays(model)
    ids2 = ids.clone()
    ids2[ids2 == 1] = 2
    ids2[1, 40:] = 1                                   # right padding with pad_token_id = 1
    with torch.no_grad():
        o = model(input_ids=ids2, attention_mask=am, labels=labels)
    arrs.update(input_ids=ids2.numpy(), attention_mask=am.numpy(), labels=labels.numpy(), logits=o.logits.numpy().copy(),
                loss=np.array(float(o.loss)))
    np.savez_compressed(os.path.join(GOLD, "opt_tiny_cls.npz"), **arrs)
    print("opt_tiny_cls", o.logits.tolist(), float(o.loss))

    torch.manual_seed(0)
    cfg = m.llama_cfg.LlamaQuantizedConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                                           vocab_size=512, max_position_embeddings=64, quant_config=deepcopy(cfgs["bfp_6bit.toml"]),
                                           num_labels=2, pad_token_id=0)
    model = m.llama.LlamaQuantizedForSequenceClassification(cfg).eval()
    arrs = {k: v for k, v in sd_arrays(model).items() if "rotary_emb" not in k}
    with torch.no_grad():
        o = model(input_ids=ids, attention_mask=am, labels=labels)
    arrs.update(input_ids=ids.numpy(), attention_mask=am.numpy(), labels=labels.numpy(), logits=o.logits.numpy().copy(),
                loss=np.array(float(o.loss)))
    np.savez_compressed(os.path.join(GOLD, "llama_tiny_cls.npz"), **arrs)
    print("llama_tiny_cls", o.logits.tolist(), float(o.loss))


if __name__ == "__main__":
    main()
