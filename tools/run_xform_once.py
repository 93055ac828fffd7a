"""One call of each quantized-Linear variant (two-launch, fused prologue, packed weights) at 4096^3 and at the decode shape
M=64, K=4096, N=16384 — for `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum` (HBM bytes per launch)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llm_mixed_q_b200 import _lib as L
from llm_mixed_q_b200.models.quantize import get_quantized_cls
from llm_mixed_q_b200.models.quantize.quantized_modules import linear as lin_mod
L.load(); dev = torch.device("cuda:0")
c = {"name": "block_fp", "bypass": False, "is_ptq": True}
for p in ("data_in", "weight", "bias"):
    c.update({f"{p}_width": 6, f"{p}_exponent_width": 8, f"{p}_exponent_bias": 127, f"{p}_block_size": [16] if p == "bias" else [1, 16]})
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for (M, K, N) in [(4096, 4096, 4096), (64, 4096, 16384)]:
    torch.manual_seed(0)
    lin = get_quantized_cls("linear", c)(K, N, bias=True, config=c).to(dev).eval()
    with torch.no_grad():
        lin.weight.normal_(0, 0.02)
        x = torch.randn(M, K, device=dev)
        lin(x)                                    # PTQ + caches
        lin_mod.PACKED_WEIGHTS = True; lin(x); lin_mod.PACKED_WEIGHTS = False
        torch.cuda.synchronize()
        for mode in ("two_launch", "fused", "packed"):
            flush.zero_()                          # evict L2 (ncu: look for the launches after each 'vectorized_elementwise' fill)
            lin_mod.FUSED_PROLOGUE, lin_mod.PACKED_WEIGHTS = mode == "fused", mode == "packed"
            torch.cuda.nvtx.range_push(f"{mode}_{M}x{K}x{N}")
            lin(x)
            torch.cuda.nvtx.range_pop()
            lin_mod.FUSED_PROLOGUE = lin_mod.PACKED_WEIGHTS = False
torch.cuda.synchronize()
