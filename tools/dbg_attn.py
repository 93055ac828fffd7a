import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llm_mixed_q_b200.models.quantize.quantized_functions.attention import fused_causal_attention
from llm_mixed_q_b200.models.quantize.quantizers import block_fp_quantizer
CFG = {"name": "block_fp", "bypass": False, "is_ptq": True}
for p in ("data_in", "weight", "bias"):
    CFG.update({f"{p}_width": 6, f"{p}_exponent_width": 8, f"{p}_exponent_bias": 127, f"{p}_block_size": [16] if p == "bias" else [1, 16]})
for d, S in [(64, 384), (64, 128), (64, 2048), (128, 384)]:
    g = torch.Generator(device="cuda").manual_seed(5 + d)
    B, heads = 2, 2
    H = heads * d
    q = torch.randn(B, S, H, device="cuda", generator=g); k = torch.randn(B, S, H, device="cuda", generator=g); v = torch.randn(B, S, H, device="cuda", generator=g)
    a = fused_causal_attention(q, k, v, CFG, CFG, heads)
    b = fused_causal_attention(q, k, v, CFG, CFG, heads)
    print(d, S, "determinism mismatches:", int((a != b).sum()), "of", a.numel())
    oq = fused_causal_attention(q, k, v, CFG, CFG, heads, out_cfg=CFG).float()
    oq2 = fused_causal_attention(q, k, v, CFG, CFG, heads, out_cfg=CFG).float()
    print("   q-out determinism mismatches:", int((oq != oq2).sum()))
    want = block_fp_quantizer(a, 6, 8, 127, [1, 16], True)
    bad = (oq != want)
    print("   oq vs quant(o32) mismatches:", int(bad.sum()))
    if bad.any():
        idx = bad.nonzero()[:8]
        for i in idx:
            i = tuple(int(t) for t in i)
            blk = (i[0], i[1], slice(i[2] // 16 * 16, i[2] // 16 * 16 + 16))
            print("    at", i, "oq", float(oq[i]), "want", float(want[i]), "o32", float(a[i]), "blockmax", float(a[blk].abs().max()))
