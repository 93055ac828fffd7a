"""A few launches of the split-attention path (block_log: rope_quantize_split, QK^T / PV plane GEMMs, softmax + P-quantizer) at the
Llama-7B layer shape, for ncu captures and quick timing.  usage: python tools/run_split_attention_once.py [width=4] [time]"""
import json, math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llm_mixed_q_b200 import _lib as L
from llm_mixed_q_b200.models.quantize.quantized_functions.split_attention import rope_quantize_split, split_attention
lib = L.load(); dev = torch.device("cuda:0")
width = int(sys.argv[1]) if len(sys.argv) > 1 else 4
cfg = {"name": "block_log", "bypass": False, "is_ptq": True}
for p in ("data_in", "weight", "bias"):
    cfg.update({f"{p}_width": width, f"{p}_exponent_bias_width": 8, f"{p}_block_size": [16] if p == "bias" else [1, 16]})
B, heads, S, d = 2, 32, 2048, 128
g = torch.Generator(device=dev).manual_seed(0)
q = torch.randn(B, S, heads * d, device=dev, generator=g) * 0.3
k = torch.randn(B, S, heads * d, device=dev, generator=g) * 0.9
v = torch.randn(B, S, heads * d, device=dev, generator=g)
def run():
    Qq, Kp = rope_quantize_split(q, k, None, None, None, None, cfg, heads)
    return split_attention(Qq, Kp, v, cfg, heads, math.sqrt(d), causal=True)
for _ in range(2):
    run()
torch.cuda.synchronize()
if len(sys.argv) > 3:
    lib.bq_set_softmax_smem_rows(int(sys.argv[3]))
if len(sys.argv) > 2:
    L.profile_enable(True)
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    prof = L.profile_read()
    L.profile_enable(False)
    print(json.dumps({k: [round(v[0] / 5, 4), v[1] // 5] for k, v in prof.items() if v[1]}, indent=1))
