import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llm_mixed_q_b200.models.quantize.quantized_functions.attention import fused_causal_attention_q
dev = torch.device("cuda:0")
cfg = {"name": "block_fp", "bypass": False, "is_ptq": True}
for p in ("data_in", "weight", "bias"):
    cfg.update({f"{p}_width": 6, f"{p}_exponent_width": 8, f"{p}_exponent_bias": 127, f"{p}_block_size": [16] if p == "bias" else [1, 16]})
B, heads, S, d = 8, 32, 2048, 64
H = heads * d
q = (torch.randn(B, S, H, device=dev) * 0.1).to(torch.bfloat16); k = torch.randn(B, S, H, device=dev).to(torch.bfloat16); v = torch.randn(B, S, H, device=dev).to(torch.bfloat16)
for _ in range(3): fused_causal_attention_q(q, k, v, cfg, heads, B, S, 1.0, out_cfg=cfg)
torch.cuda.synchronize()
