"""
ORACLE — TEST INFRASTRUCTURE ONLY.  Goldens for the fused attention kernel's masked modes, from the UNMODIFIED reference (CPU):

  tests/golden/bert_small_bfp6.npz        BertQuantizedForSequenceClassification, head_dim 64 (hidden 128, 2 heads, 2 layers), a batch
                                          with one right-padded sequence: bidirectional attention + key-padding mask
                                          (reference bert_quantized/modeling_bert.py:366-435)
  tests/golden/opt_small_padded_bfp6.npz  OPTQuantizedForCausalLM, head_dim 64, a batch with one right-padded sequence: causal mask +
                                          key padding (reference opt_quantized/modeling_opt.py:520-548, :246-312)

Usage (authoring container only):  python oracle/gen_golden_masked_attention.py
"""
import json
import os
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


def sd_arrays(model):
    return {"sd::" + k: v.detach().cpu().numpy().copy() for k, v in model.state_dict().items() if "rotary_emb" not in k}


def main():
    m = ref_loader.load_bert()
    qc = toml.load(os.path.join(REF_CFG_DIR, "bfp_6bit.toml"))
    rs = np.random.RandomState(11)
    S = 80
    ids = torch.from_numpy(rs.randint(2, 512, size=(3, S)).astype(np.int64))
    am = torch.ones(3, S, dtype=torch.long)
    am[1, 50:] = 0
    am[2, 33:] = 0
    tt = torch.zeros(3, S, dtype=torch.long)
    tt[:, 40:] = 1

    torch.manual_seed(0)
    cfg = m.bert_cfg.BertQuantizedConfig(quant_config=deepcopy(qc), initializer_range=0.05, is_decoder=False, add_cross_attention=False,
                                         chunk_size_feed_forward=0, vocab_size=512, hidden_size=128, num_hidden_layers=2,
                                         num_attention_heads=2, intermediate_size=256, max_position_embeddings=128, num_labels=3)
    model = m.bert.BertQuantizedForSequenceClassification(cfg).eval()
    arrs = sd_arrays(model)
    ids_b = ids.clone()
    ids_b[am == 0] = 0
    with torch.no_grad():
        o = model(input_ids=ids_b, attention_mask=am, token_type_ids=tt, output_hidden_states=True)
    arrs.update(input_ids=ids_b.numpy(), attention_mask=am.numpy(), token_type_ids=tt.numpy(), logits=o.logits.numpy().copy(),
                last_hidden=o.hidden_states[-1].numpy().copy())
    np.savez_compressed(os.path.join(GOLD, "bert_small_bfp6.npz"), **arrs)
    print("bert_small_bfp6", o.logits.flatten().tolist())

    torch.manual_seed(0)
    cfg = m.opt_cfg.OPTQuantizedConfig(hidden_size=128, num_hidden_layers=2, ffn_dim=256, num_attention_heads=2, vocab_size=512,
                                       max_position_embeddings=128, quant_config=deepcopy(qc), pad_token_id=1)
    model = m.opt.OPTQuantizedForCausalLM(cfg).eval()
    arrs = sd_arrays(model)
    ids_o = ids.clone()
    ids_o[am == 0] = 1
    labels = ids_o.clone()
    labels[am == 0] = -100
    with torch.no_grad():
        o = model(input_ids=ids_o, attention_mask=am, labels=labels)
    arrs.update(input_ids=ids_o.numpy(), attention_mask=am.numpy(), labels=labels.numpy(), logits=o.logits.numpy().copy(),
                loss=np.array(float(o.loss)))
    np.savez_compressed(os.path.join(GOLD, "opt_small_padded_bfp6.npz"), **arrs)
    print("opt_small_padded_bfp6", float(o.loss))


if __name__ == "__main__":
    main()
