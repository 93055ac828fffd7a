"""First-contact GPU probe: answers the on-device checklist of SURVEY.md §7 and smoke-tests the kernels.
Run on a B200 via gpurun; writes gpurun_out/probe.json."""
import ctypes, json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llm_mixed_q_b200 import _lib as L
from oracle import oracle as O

out = {}
dev = torch.device("cuda:0")
out["gpu"] = torch.cuda.get_device_name(0)
lib = L.load()

# 1. 2**e exact on CUDA for integer e in [-160, 130]?
e = torch.arange(-160, 131, dtype=torch.float32, device=dev)
p_gpu = (2 ** e)
p_ref = torch.tensor([float(2.0 ** int(k)) if -149 <= k <= 127 else (float("inf") if k > 127 else 0.0) for k in range(-160, 131)], dtype=torch.float64).to(torch.float32).to(dev)
out["pow2_exact_on_cuda"] = bool(torch.equal(p_gpu.view(torch.int32), p_ref.view(torch.int32)))
out["pow2_mismatch_e"] = e[(p_gpu.view(torch.int32) != p_ref.view(torch.int32))].tolist()[:20]

# 2. torch CPU log2 vs CUDA log2 on cliffs + random
g = torch.Generator().manual_seed(0)
bits = torch.randint(0x00000001, 0x7f800000, (1 << 24,), generator=g, dtype=torch.int32)
x = bits.view(torch.float32)
lc = torch.log2(x); lg = torch.log2(x.to(dev)).cpu()
out["log2_cpu_vs_cuda_mismatch_of_16M"] = int((lc.view(torch.int32) != lg.view(torch.int32)).sum())
out["ceil_log2_cpu_vs_cuda_mismatch_of_16M"] = int((torch.ceil(lc) != torch.ceil(lg)).sum())
out["floor_log2_cpu_vs_cuda_mismatch_of_16M"] = int((torch.floor(lc) != torch.floor(lg)).sum())
out["round_log2_cpu_vs_cuda_mismatch_of_16M"] = int((torch.round(lc) != torch.round(lg)).sum())

def fmt(kind, width=0, ew=0, bias=0, bw=0, br=1, bc=16, fold=0):
    return L.BqFormat(L.KIND[kind], width, ew, bias, bw, br, bc, fold)

def quant(f, x3, dtype=torch.float32, transpose=False):
    Lz, R, C = x3.shape
    t = L.BqTensor3(Lz, R, C, x3.stride(0), x3.stride(1), x3.stride(2))
    y = torch.empty((Lz, C, R) if transpose else (Lz, R, C), dtype=dtype, device=dev)
    n = lib.bq_quantize_workspace_bytes(ctypes.byref(f), ctypes.byref(t))
    ws = torch.empty(max(n, 256), dtype=torch.uint8, device=dev)
    rc = lib.bq_quantize(ctypes.byref(f), ctypes.byref(t), x3.data_ptr(), y.data_ptr(), 0 if dtype == torch.float32 else 1,
                         1 if transpose else 0, ws.data_ptr(), ws.numel(), L.stream_ptr())
    L.check(rc, "bq_quantize")
    return y

def eqbits(a, b):
    return int((a.contiguous().view(torch.int32) != b.contiguous().view(torch.int32)).sum())

# 3. quantizer parity vs the oracle running ON THE GPU (same torch ops the reference would run after .to("cuda"))
torch.manual_seed(0)
res = {}
for sigma in (1e-3, 0.02, 1.0, 30.0):
    x = (torch.randn(64, 256, 1024, device=dev) * sigma)
    x.view(-1)[::13] = 0
    x[:, ::5, :32] = 0
    for name, f, ofn in [
        ("bfp6", fmt("block_fp", 6, 8, 127, fold=1), lambda t: O.block_fp_quantize(t, 6, 8, 127, [1, 16], True)),
        ("bfp4", fmt("block_fp", 4, 8, 127, fold=1), lambda t: O.block_fp_quantize(t, 4, 8, 127, [1, 16], True)),
        ("bmf8", fmt("block_minifloat", 8, 4, 0, 8, fold=1), lambda t: O.block_minifloat_quantize(t, 8, 4, 8, [1, 16], True)),
        ("bmf4", fmt("block_minifloat", 4, 2, 0, 8, fold=1), lambda t: O.block_minifloat_quantize(t, 4, 2, 8, [1, 16], True)),
        ("bl8", fmt("block_log", 8, 0, 0, 8, fold=1), lambda t: O.block_log_quantize(t, 8, 8, [1, 16], True)),
        ("bl4", fmt("block_log", 4, 0, 0, 8, fold=1), lambda t: O.block_log_quantize(t, 4, 8, [1, 16], True)),
        ("dmf8", fmt("minifloat_denorm", 8, 4, 7), lambda t: O.minifloat_denorm_quantize(t, 8, 4, 7)),
    ]:
        y = quant(f, x)
        yo = ofn(x)
        res[f"{name}_s{sigma}"] = eqbits(y, yo)
        yc = ofn(x.cpu())
        res[f"{name}_s{sigma}_cpu_oracle_vs_gpu_oracle"] = eqbits(yo.cpu(), yc)
out["quant_mismatch_vs_gpu_oracle"] = res

# 4. generic / tile paths
x = torch.randn(3, 48, 40, device=dev)
f = fmt("block_fp", 6, 8, 127, br=2, bc=16, fold=1)
out["generic_2x16"] = eqbits(quant(f, x), O.block_fp_quantize(x, 6, 8, 127, [2, 16], True))
xt = torch.randn(5, 200, 64, device=dev).transpose(1, 2)   # kT view [5, 64, 200]
f = fmt("block_fp", 6, 8, 127, fold=1)
out["tile_kT"] = eqbits(quant(f, xt), O.block_fp_quantize(xt, 6, 8, 127, [1, 16], True))
v = torch.randn(5, 200, 64, device=dev)
out["tile_transposed_out"] = eqbits(quant(f, v, transpose=True), O.block_fp_quantize(v, 6, 8, 127, [1, 16], True).transpose(1, 2).contiguous())

# 5. GEMM
def gemm(A, B, bias=None):
    bsz, M, K = A.shape
    N = B.shape[-2]
    C = torch.empty(bsz, M, N, device=dev, dtype=torch.float32)
    sb = 0 if B.dim() == 2 else B.stride(0)
    rc = lib.bq_gemm_bf16_tn(A.data_ptr(), B.data_ptr(), C.data_ptr(), bias.data_ptr() if bias is not None else None,
                             bsz, M, N, K, A.stride(1), B.stride(-2), N, A.stride(0), sb, M * N, L.stream_ptr())
    L.check(rc, "gemm")
    return C
gres = {}
for (bsz, M, N, K) in [(1, 128, 256, 64), (1, 256, 512, 256), (1, 4096, 4096, 4096), (1, 1000, 520, 328), (3, 200, 136, 64), (4, 512, 64, 512), (2, 300, 100, 1000)]:
    A = torch.randn(bsz, M, K, device=dev).to(torch.bfloat16)
    B = torch.randn(bsz, N, K, device=dev).to(torch.bfloat16) if bsz > 1 else torch.randn(N, K, device=dev).to(torch.bfloat16)
    bias = torch.randn(N, device=dev)
    try:
        C = gemm(A, B, bias)
        torch.cuda.synchronize()
        ref = torch.matmul(A.double(), (B.double().transpose(-1, -2))) + bias.double()
        err = (C.double() - ref).abs().max().item()
        gres[f"{bsz}x{M}x{N}x{K}"] = {"max_abs_err": err, "ref_absmax": ref.abs().max().item()}
    except Exception as ex:
        gres[f"{bsz}x{M}x{N}x{K}"] = {"error": repr(ex)}
out["gemm"] = gres

# 6. timing: quantizer GB/s and GEMM TFLOP/s
def timeit(fn, n=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e_ = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e_.record(); torch.cuda.synchronize()
    return s.elapsed_time(e_) / n
x = torch.randn(8, 2048, 8192, device=dev)   # 512 MiB in, 512 MiB out: larger than L2
t = L.BqTensor3(*x.shape, *x.stride())
y = torch.empty_like(x); ybf = torch.empty(x.shape, dtype=torch.bfloat16, device=dev)
tim = {}
for name, f in [("bfp6", fmt("block_fp", 6, 8, 127)), ("bmf8", fmt("block_minifloat", 8, 4, 0, 8)), ("bl8", fmt("block_log", 8, 0, 0, 8)), ("dmf8", fmt("minifloat_denorm", 8, 4, 7)), ("none", fmt("none"))]:
    n = lib.bq_quantize_workspace_bytes(ctypes.byref(f), ctypes.byref(t))
    ws = torch.empty(max(n, 256), dtype=torch.uint8, device=dev)
    ms = timeit(lambda: lib.bq_quantize(ctypes.byref(f), ctypes.byref(t), x.data_ptr(), y.data_ptr(), 0, 0, ws.data_ptr(), ws.numel(), L.stream_ptr()))
    tim[name + "_f32_GBs"] = x.numel() * 8 / ms / 1e6
    ms = timeit(lambda: lib.bq_quantize(ctypes.byref(f), ctypes.byref(t), x.data_ptr(), ybf.data_ptr(), 1, 0, ws.data_ptr(), ws.numel(), L.stream_ptr()))
    tim[name + "_bf16_GBs"] = x.numel() * 6 / ms / 1e6
ms = timeit(lambda: y.copy_(x)); tim["torch_copy_GBs"] = x.numel() * 8 / ms / 1e6
del x, y, ybf
for (M, N, K) in [(4096, 4096, 4096), (8192, 8192, 8192), (16384, 2048, 2048), (16384, 8192, 2048), (16384, 2048, 8192)]:
    A = torch.randn(1, M, K, device=dev).to(torch.bfloat16); B = torch.randn(N, K, device=dev).to(torch.bfloat16)
    C = torch.empty(1, M, N, device=dev)
    ms = timeit(lambda: lib.bq_gemm_bf16_tn(A.data_ptr(), B.data_ptr(), C.data_ptr(), None, 1, M, N, K, K, K, N, 0, 0, 0, L.stream_ptr()))
    tim[f"gemm_{M}x{N}x{K}_TFLOPs"] = 2 * M * N * K / ms / 1e9
    ms = timeit(lambda: torch.matmul(A[0], B.t()))
    tim[f"cublas_bf16_{M}x{N}x{K}_TFLOPs"] = 2 * M * N * K / ms / 1e9
out["timing"] = tim
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)
print(json.dumps(out, indent=1))
