import sys, torch, ctypes
sys.path.insert(0, '/root/repo')
from llm_mixed_q_b200 import _lib as L
from llm_mixed_q_b200.models.quantize import get_quantized_func
from llm_mixed_q_b200.models.quantize.quantized_modules.linear import quantize_operand_bf16
from oracle import oracle as O
CFG = {"name": "block_fp", "bypass": False, "is_ptq": True}
for p in ("data_in", "weight", "bias"):
    CFG.update({f"{p}_width": 6, f"{p}_exponent_width": 8, f"{p}_exponent_bias": 127, f"{p}_block_size": [16] if p == "bias" else [1, 16]})
g = torch.Generator(device="cuda").manual_seed(2)
BH, S, d = 16, 2048, 64
q = torch.randn(BH, S, d, device="cuda", generator=g); k = torch.randn(BH, S, d, device="cuda", generator=g); v = torch.randn(BH, S, d, device="cuda", generator=g)
fn = get_quantized_func("bmm", CFG)
s = fn(q, k.transpose(1, 2), config=CFG)
mask = torch.triu(torch.ones(S, S, dtype=torch.bool, device="cuda"), diagonal=1)
p = torch.softmax(s.masked_fill(mask, torch.finfo(torch.float32).min), dim=-1)
o = fn(p, v, config=CFG)
pq = O.operand_quantizer(CFG, "data_in", True)(p); vq = O.operand_quantizer(CFG, "weight", True)(v)
ex = pq.double() @ vq.double()
err = (o.double() - ex).abs()
print("max err", err.max().item(), "at", (err == err.max()).nonzero()[:3].tolist())
print("err per batch", err.amax(dim=(1, 2)).tolist())
rows = err.amax(dim=(0, 2)); print("rows with err>1e-4:", (rows > 1e-4).nonzero().flatten()[:40].tolist(), int((rows > 1e-4).sum()))
# operand checks
pbf = quantize_operand_bf16(p, "block_fp", dict(width=6, exponent_width=8, exponent_bias=127), [1, 16], True)
print("P bf16 vs oracle mismatches:", int((pbf.float() != pq).sum()), "max diff", (pbf.float() - pq).abs().max().item())
vbf = quantize_operand_bf16(v, "block_fp", dict(width=6, exponent_width=8, exponent_bias=127), [1, 16], True, transpose_out=True)
print("V bf16T vs oracle mismatches:", int((vbf.float() != vq.transpose(1, 2)).sum()))
# direct gemm with oracle operands
lib = L.load()
A = pq.to(torch.bfloat16).contiguous(); B = vq.transpose(1, 2).contiguous().to(torch.bfloat16)
C = torch.empty(BH, S, d, device="cuda")
rc = lib.bq_gemm_bf16_tn(A.data_ptr(), B.data_ptr(), C.data_ptr(), None, BH, S, d, S, S, S, d, S * S, d * S, S * d, L.stream_ptr()); L.check(rc, "g")
e2 = (C.double() - ex).abs(); print("direct gemm max err", e2.max().item(), "per batch", e2.amax(dim=(1, 2)).tolist())
A2 = torch.randn(BH, S, S, device="cuda").to(torch.bfloat16)
rc = lib.bq_gemm_bf16_tn(A2.data_ptr(), B.data_ptr(), C.data_ptr(), None, BH, S, d, S, S, S, d, S * S, d * S, S * d, L.stream_ptr())
e3 = (C.double() - A2.double() @ B.double().transpose(1, 2)).abs(); print("random A gemm max err", e3.max().item(), "per batch", e3.amax(dim=(1, 2)).tolist())
absprod = pq.double().abs() @ vq.double().abs()
bound = 4 * (S ** 0.5) * 2.0 ** -24 * absprod + 1e-30
ratio = err / bound
i = (ratio == ratio.max()).nonzero()[0].tolist()
print("worst ratio", ratio.max().item(), "at", i, "err", err[tuple(i)].item(), "absprod", absprod[tuple(i)].item(), "out", o[tuple(i)].item(), "exact", ex[tuple(i)].item())
b_, r_, c_ = i
row = pq[b_, r_]; nz = (row != 0).nonzero().flatten()
print("nnz in row", nz.numel(), "row vals", row[nz][:8].tolist(), "v", vq[b_, nz[:8], c_].tolist())
print("ratio>1 count", int((ratio > 1).sum()), "of", ratio.numel(), "; ratio>8:", int((ratio > 8).sum()))
