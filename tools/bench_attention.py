"""A/B of the fused causal attention kernels at the OPT-1.3B layer shape (B8 h32 S2048 d64): two-pipeline / P-in-TMEM kernel against
the single-pipeline / P-in-smem kernel, both numerator modes.  Prints one JSON object; also written to gpurun_out/bench_attention.json."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llm_mixed_q_b200 import _lib as L
from llm_mixed_q_b200.models.quantize.quantized_functions.attention import fused_causal_attention_q

lib = L.load(); dev = torch.device("cuda:0")
cfg = {"name": "block_fp", "bypass": False, "is_ptq": True}
for p in ("data_in", "weight", "bias"):
    cfg.update({f"{p}_width": 6, f"{p}_exponent_width": 8, f"{p}_exponent_bias": 127, f"{p}_block_size": [16] if p == "bias" else [1, 16]})


def timeit(fn, n=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n


out = {}
shapes = [(8, 32, 2048, 64), (2, 32, 2048, 64), (8, 12, 2048, 64)]
if len(sys.argv) > 1 and sys.argv[1] == "d128":
    shapes = [(2, 32, 2048, 128)]
    # Llama-7B W4A4 block_minifloat (BASELINE configs[3]): P quantised per element (shared bias per block), score / sqrt(128) applied post-matmul
    bmf = {"name": "block_minifloat", "bypass": False, "is_ptq": True}
    for p in ("data_in", "weight", "bias"):
        bmf.update({f"{p}_width": 4, f"{p}_exponent_width": 2, f"{p}_exponent_bias_width": 8, f"{p}_block_size": [16] if p == "bias" else [1, 16]})
    B, heads, S, d = shapes[0]
    g = torch.Generator(device=dev).manual_seed(0)
    q = (torch.randn(B, S, heads * d, device=dev, generator=g)).to(torch.bfloat16)
    k = (torch.randn(B, S, heads * d, device=dev, generator=g)).to(torch.bfloat16)
    v = torch.randn(B, S, heads * d, device=dev, generator=g).to(torch.bfloat16)
    out[f"B{B}h{heads}S{S}d{d}_block_minifloat_w4_ms"] = round(timeit(lambda: fused_causal_attention_q(q, k, v, bmf, heads, B, S, 128 ** 0.5, out_cfg=bmf), n=20), 4)
if len(sys.argv) > 1 and sys.argv[1] == "peaked":
    # peaked softmax rows (trained checkpoints): score std 0.8 (random init) / 4 / 8 -> share of pass-through probabilities (<= 1e-8,
    # returned unquantised by the reference) ~0 / most / nearly all; default kernel only
    B, heads, S, d = 8, 32, 2048, 64
    for q_std in (0.113, 0.55, 1.1):
        g = torch.Generator(device=dev).manual_seed(0)
        q = (torch.randn(B, S, heads * d, device=dev, generator=g) * q_std).to(torch.bfloat16)
        k = (torch.randn(B, S, heads * d, device=dev, generator=g) * 0.9).to(torch.bfloat16)
        v = torch.randn(B, S, heads * d, device=dev, generator=g).to(torch.bfloat16)
        ms = timeit(lambda: fused_causal_attention_q(q, k, v, cfg, heads, B, S, 1.0, out_cfg=cfg), n=20)
        out[f"B{B}h{heads}S{S}d{d}_score_std_{q_std * 0.9 * 8:.1f}_ms"] = round(ms, 4)
    shapes = []
    json.dump(out, open("gpurun_out/bench_attention_peaked.json", "w"), indent=1)
for (B, heads, S, d) in shapes:
    Hh = heads * d
    g = torch.Generator(device=dev).manual_seed(0)
    q = (torch.randn(B, S, Hh, device=dev, generator=g) * 0.113).to(torch.bfloat16)      # random-init OPT statistics: scores ~ N(0, 0.8)
    k = (torch.randn(B, S, Hh, device=dev, generator=g) * 0.9).to(torch.bfloat16)
    v = torch.randn(B, S, Hh, device=dev, generator=g).to(torch.bfloat16)
    for dual in (1, 0):
        for precise in (0, 1):
            lib.bq_set_attention_dual_pipeline(dual); lib.bq_set_attention_precise_exp(precise)
            ms = timeit(lambda: fused_causal_attention_q(q, k, v, cfg, heads, B, S, 1.0, out_cfg=cfg), n=20)
            key = f"B{B}h{heads}S{S}d{d}_{'dual' if dual else 'single'}_{'expf' if precise else 'ex2'}"
            out[key + "_ms"] = round(ms, 4)
            out[key + "_Gscores_per_s"] = round(B * heads * S * (S + 1) / 2 / ms / 1e6, 1)
lib.bq_set_attention_dual_pipeline(1); lib.bq_set_attention_precise_exp(0)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/bench_attention.json", "w"), indent=1)
print(json.dumps(out, indent=1))
