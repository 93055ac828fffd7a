"""
ORACLE — TEST INFRASTRUCTURE ONLY.  Golden for the block_log attention path (split_attention.py) from the UNMODIFIED reference (CPU):

  tests/golden/llama_small_bl8.npz   LlamaQuantizedForCausalLM under block_log.toml, head_dim 64 (hidden 128, 2 heads, 2 layers),
                                     batch of 3 x 96 tokens, one sequence right-padded: matmul_0 / matmul_1 quantise x only and keep
                                     k / v in fp32 (reference quantized_functions/matmul.py:286-297), causal + key-padding mask
                                     (models/llama_quantized/modeling_llama.py:309-337)
  tests/golden/opt_small_bl8.npz     OPTQuantizedForCausalLM under block_log.toml, head_dim 64, same token batch (right-padded row)

Usage (authoring container only):  python oracle/gen_golden_llama_block_log.py
"""
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


def main():
    m = ref_loader.load_models()
    qc = toml.load(os.path.join(REF_CFG_DIR, "block_log.toml"))
    rs = np.random.RandomState(21)
    S = 96
    ids = torch.from_numpy(rs.randint(2, 512, size=(3, S)).astype(np.int64))
    am = torch.ones(3, S, dtype=torch.long)
    am[1, 70:] = 0
    torch.manual_seed(0)
    cfg = m.llama_cfg.LlamaQuantizedConfig(hidden_size=128, intermediate_size=352, num_hidden_layers=2, num_attention_heads=2,
                                           vocab_size=512, max_position_embeddings=128, initializer_range=0.05, pad_token_id=0,
                                           quant_config=deepcopy(qc))
    model = m.llama.LlamaQuantizedForCausalLM(cfg).eval()
    arrs = {"sd::" + k: v.detach().cpu().numpy().copy() for k, v in model.state_dict().items() if "rotary_emb" not in k}
    ids[am == 0] = 0
    labels = ids.clone()
    labels[am == 0] = -100
    with torch.no_grad():
        o = model(input_ids=ids, attention_mask=am, labels=labels)
        o2 = model(input_ids=ids[:1], labels=ids[:1])
    arrs.update(input_ids=ids.numpy(), attention_mask=am.numpy(), labels=labels.numpy(), logits=o.logits.numpy().copy(),
                loss=np.array(float(o.loss)), logits_unpadded_row0=o2.logits.numpy().copy(), loss_unpadded_row0=np.array(float(o2.loss)))
    np.savez_compressed(os.path.join(GOLD, "llama_small_bl8.npz"), **arrs)
    print("llama_small_bl8", float(o.loss), float(o2.loss))

    # OPT under block_log.toml, head_dim 64: bmm_0 / bmm_1 quantise x only (q is scaled before bmm_0, modeling_opt.py:206-246)
    torch.manual_seed(0)
    ocfg = m.opt_cfg.OPTQuantizedConfig(hidden_size=128, num_hidden_layers=2, ffn_dim=256, num_attention_heads=2, vocab_size=512,
                                        max_position_embeddings=128, quant_config=deepcopy(qc), pad_token_id=1, init_std=0.05)
    omodel = m.opt.OPTQuantizedForCausalLM(ocfg).eval()
    arrs = {"sd::" + k: v.detach().cpu().numpy().copy() for k, v in omodel.state_dict().items()}
    oids = ids.clone()
    oids[am == 0] = 1
    olabels = oids.clone()
    olabels[am == 0] = -100
    with torch.no_grad():
        o = omodel(input_ids=oids, attention_mask=am, labels=olabels)
    arrs.update(input_ids=oids.numpy(), attention_mask=am.numpy(), labels=olabels.numpy(), logits=o.logits.numpy().copy(),
                loss=np.array(float(o.loss)))
    np.savez_compressed(os.path.join(GOLD, "opt_small_bl8.npz"), **arrs)
    print("opt_small_bl8", float(o.loss))


if __name__ == "__main__":
    main()
