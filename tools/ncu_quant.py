"""ncu target: two launches of the streaming quantizer per format (fp32 out) on a [8,2048,8192] tensor.
usage: ncu --set full --clock-control none --import-source on -k regex:quant_rows -o gpurun_out/quant python tools/ncu_quant.py"""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llm_mixed_q_b200 import _lib as L
lib = L.load(); dev = torch.device("cuda:0")
def fmt(kind, width=0, ew=0, bias=0, bw=0, br=1, bc=16, fold=0):
    return L.BqFormat(L.KIND[kind], width, ew, bias, bw, br, bc, fold)
x = torch.randn(8, 2048, 8192, device=dev); x.view(-1)[::13] = 0
t = L.BqTensor3(*x.shape, *x.stride())
y = torch.empty_like(x)
which = sys.argv[1:] or ["bfp6", "bmf8", "bl8", "dmf8"]
F = {"bfp6": fmt("block_fp", 6, 8, 127), "bmf8": fmt("block_minifloat", 8, 4, 0, 8), "bmf4": fmt("block_minifloat", 4, 2, 0, 8),
     "bl8": fmt("block_log", 8, 0, 0, 8), "bl4": fmt("block_log", 4, 0, 0, 8), "dmf8": fmt("minifloat_denorm", 8, 4, 7)}
for name in which:
    f = F[name]
    n = lib.bq_quantize_workspace_bytes(ctypes.byref(f), ctypes.byref(t))
    ws = torch.empty(max(n, 256), dtype=torch.uint8, device=dev)
    for _ in range(2):
        rc = lib.bq_quantize(ctypes.byref(f), ctypes.byref(t), x.data_ptr(), y.data_ptr(), 0, 0, ws.data_ptr(), ws.numel(), L.stream_ptr())
        assert rc == 0
    torch.cuda.synchronize()
print("done")
