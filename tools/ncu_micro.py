"""One launch of each fused kernel at the OPT-1.3B layer shapes (target for `ncu --set full`)."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llm_mixed_q_b200 import _lib as L
from llm_mixed_q_b200.models.quantize.quantized_functions.fused_glue import norm_quantize
from llm_mixed_q_b200.models.quantize.quantizers.utils import make_format
lib = L.load(); dev = torch.device("cuda:0")
f6 = ("block_fp", dict(width=6, exponent_width=8, exponent_bias=127))
x = torch.randn(16384, 2048, device=dev); w = torch.ones(2048, device=dev); b = torch.zeros(2048, device=dev)
for _ in range(2): norm_quantize(x, w, b, 1e-5, [f6])
fq = make_format("block_fp", width=6, exponent_width=8, exponent_bias=127, b0=1, b1=16)
M, N, K = 16384, 2048, 2048
A = torch.randn(M, K, device=dev).to(torch.bfloat16); Bw = (torch.randn(N, K, device=dev) * 0.02).to(torch.bfloat16)
bias = torch.randn(N, device=dev) * 0.02; res = torch.randn(M, N, device=dev)
for mode in ("q_n", "q_m", "residual"):
    ep = L.BqGemmEpilogue(); ep.bias = bias.data_ptr(); ep.scale = 1.0
    bf = mode != "residual"
    if mode == "residual": ep.residual, ep.ldr = res.data_ptr(), N
    else: ep.qfmt = C.pointer(fq); ep.qdir = 1 if mode == "q_m" else 0
    ep.out_dtype = 1 if bf else 0
    Cc = torch.empty(M, N, device=dev, dtype=torch.bfloat16 if bf else torch.float32)
    for _ in range(2):
        lib.bq_gemm_bf16_tn_ex(A.data_ptr(), Bw.data_ptr(), Cc.data_ptr(), C.byref(ep), M, N, K, K, K, N, L.stream_ptr())
torch.cuda.synchronize()
