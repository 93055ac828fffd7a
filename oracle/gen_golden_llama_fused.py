"""
ORACLE — TEST INFRASTRUCTURE ONLY.  Goldens for the fused Llama layer (RMSNorm + quantize, q | k | v in one GEMM with the RoPE /
quantizer epilogue, one-kernel attention at head_dim 128, gate | up in one GEMM with the gated-SiLU epilogue, residual epilogues) from the
UNMODIFIED reference run on the CPU in this container:

  tests/golden/llama_small_bfp6.npz   LlamaQuantizedForCausalLM under bfp_6bit.toml, hidden 256 = 2 heads x 128, intermediate 352,
                                      2 layers, batch of 3 x 96 tokens, one sequence right-padded; also row 0 alone, unpadded
  tests/golden/llama_small_bmf8.npz   same under block_minifloat.toml (weights drawn with std 1.5: with the usual 0.02 every weight
                                      block has max < 1 and block_minifloat maps it to 0)

Reference code path: models/llama_quantized/modeling_llama.py:246 (MLP), :274-344 (attention), quantized_functions/matmul.py:146-297,
quantized_functions/rotary_positional_encoding.py:27-36, quantized_modules/linear.py:59-76.

Usage (authoring container only):  python oracle/gen_golden_llama_fused.py
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
    rs = np.random.RandomState(33)
    S = 96
    ids0 = torch.from_numpy(rs.randint(2, 512, size=(3, S)).astype(np.int64))
    am = torch.ones(3, S, dtype=torch.long)
    am[2, 80:] = 0
    for tag, tomlname, init in (("llama_small_bfp6", "bfp_6bit.toml", 0.05), ("llama_small_bmf8", "block_minifloat.toml", 1.5)):
        qc = toml.load(os.path.join(REF_CFG_DIR, tomlname))
        torch.manual_seed(0)
        cfg = m.llama_cfg.LlamaQuantizedConfig(hidden_size=256, intermediate_size=352, num_hidden_layers=2, num_attention_heads=2,
                                               vocab_size=512, max_position_embeddings=128, initializer_range=init, pad_token_id=0,
                                               quant_config=deepcopy(qc))
        model = m.llama.LlamaQuantizedForCausalLM(cfg).eval()
        arrs = {"sd::" + k: v.detach().cpu().numpy().copy() for k, v in model.state_dict().items() if "rotary_emb" not in k}
        ids = ids0.clone()
        ids[am == 0] = 0
        labels = ids.clone()
        labels[am == 0] = -100
        with torch.no_grad():
            o = model(input_ids=ids, attention_mask=am, labels=labels)
            o2 = model(input_ids=ids[:1], labels=ids[:1])
        arrs.update(input_ids=ids.numpy(), attention_mask=am.numpy(), labels=labels.numpy(), logits=o.logits.numpy().copy(),
                    loss=np.array(float(o.loss)), logits_unpadded_row0=o2.logits.numpy().copy(), loss_unpadded_row0=np.array(float(o2.loss)),
                    initializer_range=np.array(init))
        np.savez_compressed(os.path.join(GOLD, tag + ".npz"), **arrs)
        print(tag, float(o.loss), float(o2.loss), float(o.logits.std()))


if __name__ == "__main__":
    main()
