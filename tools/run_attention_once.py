"""One launch of each attention kernel variant at the OPT-1.3B layer shape (for ncu captures)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llm_mixed_q_b200 import _lib as L
from llm_mixed_q_b200.models.quantize.quantized_functions.attention import fused_causal_attention_q
lib = L.load(); dev = torch.device("cuda:0")
cfg = {"name": "block_fp", "bypass": False, "is_ptq": True}
for p in ("data_in", "weight", "bias"):
    cfg.update({f"{p}_width": 6, f"{p}_exponent_width": 8, f"{p}_exponent_bias": 127, f"{p}_block_size": [16] if p == "bias" else [1, 16]})
B, heads, S, d = 8, 32, 2048, int(sys.argv[1]) if len(sys.argv) > 1 else 64
if d == 128: B = 2
g = torch.Generator(device=dev).manual_seed(0)
q = (torch.randn(B, S, heads * d, device=dev, generator=g) * 0.113).to(torch.bfloat16)   # random-init OPT statistics: scores ~ N(0, 0.8)
k = (torch.randn(B, S, heads * d, device=dev, generator=g) * 0.9).to(torch.bfloat16)
v = torch.randn(B, S, heads * d, device=dev, generator=g).to(torch.bfloat16)
for dual in (1, 0):
    lib.bq_set_attention_dual_pipeline(dual)
    for _ in range(2):
        fused_causal_attention_q(q, k, v, cfg, heads, B, S, 1.0, out_cfg=cfg)
torch.cuda.synchronize()
