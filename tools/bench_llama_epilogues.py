"""A/B of the two Llama GEMM-epilogue fusions at the Llama-7B layer shape (4096 tokens, H 4096, I 11008, head_dim 128), same process, same box:
  q_proj / k_proj:  fp32 GEMM + bq_rope_quantize  vs  bq_gemm_bf16_tn_rope (RoPE + matmul_0 operand quantizer in the epilogue)
  MLP operand:      gate GEMM + up GEMM + silu*mul quantizer  vs  one GEMM over interleaved weights with the gated epilogue (act = 2)
Prints one JSON object; also written to gpurun_out/bench_llama_epilogues.json."""
import copy, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llm_mixed_q_b200 import _lib as L
from llm_mixed_q_b200.models.llama_quantized.modeling_llama import LlamaRotaryEmbedding
from llm_mixed_q_b200.models.quantize import get_quantized_cls
from llm_mixed_q_b200.models.quantize.quantized_functions.fused_glue import silu_mul_quantize
from llm_mixed_q_b200.models.quantize.quantized_functions.rotary_positional_encoding import apply_token_major_quantized, rope_quantize_operands
from llm_mixed_q_b200.models.quantize.quantized_modules import linear as QL

dev = torch.device("cuda:0")
cfg = {"name": "block_minifloat", "bypass": False, "is_ptq": True}
for p in ("data_in", "weight", "bias"):
    cfg.update({f"{p}_width": 4, f"{p}_exponent_width": 2, f"{p}_exponent_bias_width": 8, f"{p}_block_size": [16] if p == "bias" else [1, 16]})
m0 = {k: v for k, v in cfg.items() if not k.startswith("bias")}
rope_cfg = {"name": "integer", "bypass": False, "data_in_width": 8, "data_in_frac_width": 7}
fmt = ("block_minifloat", dict(width=4, exponent_width=2, exponent_bias_width=8))


def timeit(fn, n=30, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n


B, S, H, I, heads, d = 2, 2048, 4096, 11008, 32, 128
M = B * S
mk = lambda k, n: get_quantized_cls("linear", cfg)(k, n, bias=False, config=copy.deepcopy(cfg)).to(dev)
lq, lk, lg, lu = mk(H, H), mk(H, H), mk(H, I), mk(H, I)
with torch.no_grad():
    for m in (lq, lk, lg, lu):
        m.weight.mul_(64.0)
g = torch.Generator(device=dev).manual_seed(0)
xq = QL.quantize_operand_bf16(torch.randn(M, H, device=dev, generator=g), fmt[0], fmt[1], [1, 16], True)
# a second, different operand set so that consecutive timed calls do not find everything in L2
rot = LlamaRotaryEmbedding(d, max_position_embeddings=S).to(dev)
cos, sin = rot(torch.zeros(1, device=dev), seq_len=S)
cos_t, sin_t, pos, fq, fk = rope_quantize_operands(cos, sin, None, rope_cfg, m0, B, S, d)
out = {}


def qk_unfused():
    q, k = lq.forward_prequantized(xq), lk.forward_prequantized(xq)
    return apply_token_major_quantized(q.view(B, S, H), k.view(B, S, H), cos, sin, None, rope_cfg, m0, heads)


def qk_fused():
    return (QL.rope_prequantized(lq, xq, cos_t, sin_t, pos, fq, S, d, False), QL.rope_prequantized(lk, xq, cos_t, sin_t, pos, fk, S, d, True))


lv = mk(H, H)
with torch.no_grad():
    lv.weight.mul_(64.0)


def qkv_three():
    return qk_fused() + (lv.forward_prequantized(xq, out_format=fmt),)


def qkv_one():
    return QL.qkv_rope_prequantized(lq, lk, lv, xq, cos_t, sin_t, pos, fq, fk, fmt, S, d)


a3 = qkv_three()
out["qkv_three_launches_ms"] = round(timeit(qkv_three), 4)
b3 = qkv_one()
out["qkv_one_launch_bit_identical"] = bool(all(torch.equal(x, y) for x, y in zip(a3, b3)))
out["qkv_one_launch_ms"] = round(timeit(qkv_one), 4)
a, b = qk_unfused(), qk_fused()
out["qk_bit_identical"] = bool(torch.equal(a[0].view(M, H), b[0]) and torch.equal(a[1].view(M, H), b[1]))
out["q_gemm_fp32_ms"] = round(timeit(lambda: lq.forward_prequantized(xq)), 4)
out["q_rope_gemm_ms"] = round(timeit(lambda: QL.rope_prequantized(lq, xq, cos_t, sin_t, pos, fq, S, d, False)), 4)
out["k_rope_gemm_ms"] = round(timeit(lambda: QL.rope_prequantized(lk, xq, cos_t, sin_t, pos, fk, S, d, True)), 4)
out["qk_unfused_ms"] = round(timeit(qk_unfused), 4)
out["qk_fused_ms"] = round(timeit(qk_fused), 4)


def mlp_unfused():
    return silu_mul_quantize(lg.forward_prequantized(xq), lu.forward_prequantized(xq), fmt)


def mlp_fused():
    return QL.gated_silu_prequantized(lg, lu, xq, fmt)


a = mlp_unfused()
out["mlp_unfused_ms"] = round(timeit(mlp_unfused), 4)
b = mlp_fused()
out["mlp_bit_identical"] = bool(torch.equal(a, b))
out["mlp_fused_ms"] = round(timeit(mlp_fused), 4)
out["shape"] = f"tokens {M}, H {H}, I {I}, heads {heads} x {d}; W4A4 block_minifloat"
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/bench_llama_epilogues.json", "w"), indent=1)
print(json.dumps(out, indent=1))
