"""Kernel micro-benchmarks on one B200 (not the headline bench): quantizer GB/s per format and GEMM TFLOP/s per shape."""
import ctypes, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llm_mixed_q_b200 import _lib as L
lib = L.load(); dev = torch.device("cuda:0")
def fmt(kind, width=0, ew=0, bias=0, bw=0, br=1, bc=16, fold=0):
    return L.BqFormat(L.KIND[kind], width, ew, bias, bw, br, bc, fold)
def timeit(fn, n=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e_ = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e_.record(); torch.cuda.synchronize()
    return s.elapsed_time(e_) / n
out = {}
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "quant"):
    x = torch.randn(8, 2048, 8192, device=dev); x.view(-1)[::13] = 0
    t = L.BqTensor3(*x.shape, *x.stride())
    y = torch.empty_like(x); ybf = torch.empty(x.shape, dtype=torch.bfloat16, device=dev)
    for name, f in [("bfp6", fmt("block_fp", 6, 8, 127)), ("bfp4", fmt("block_fp", 4, 8, 127)), ("bmf8", fmt("block_minifloat", 8, 4, 0, 8)), ("bmf4", fmt("block_minifloat", 4, 2, 0, 8)),
                    ("bl8", fmt("block_log", 8, 0, 0, 8)), ("bl4", fmt("block_log", 4, 0, 0, 8)), ("dmf8", fmt("minifloat_denorm", 8, 4, 7)), ("none", fmt("none"))]:
        n = lib.bq_quantize_workspace_bytes(ctypes.byref(f), ctypes.byref(t))
        ws = torch.empty(max(n, 256), dtype=torch.uint8, device=dev)
        ms = timeit(lambda: lib.bq_quantize(ctypes.byref(f), ctypes.byref(t), x.data_ptr(), y.data_ptr(), 0, 0, ws.data_ptr(), ws.numel(), L.stream_ptr()))
        out[name + "_f32_GBs"] = round(x.numel() * 8 / ms / 1e6, 1)
        ms = timeit(lambda: lib.bq_quantize(ctypes.byref(f), ctypes.byref(t), x.data_ptr(), ybf.data_ptr(), 1, 0, ws.data_ptr(), ws.numel(), L.stream_ptr()))
        out[name + "_bf16_GBs"] = round(x.numel() * 6 / ms / 1e6, 1)
    ms = timeit(lambda: y.copy_(x)); out["torch_copy_GBs"] = round(x.numel() * 8 / ms / 1e6, 1)
    # softmax-probability input (half zeros under the causal mask)
    s = torch.randn(64, 2048, 2048, device=dev) * 3
    mask = torch.triu(torch.ones(2048, 2048, dtype=torch.bool, device=dev), diagonal=1)
    p = torch.softmax(s.masked_fill(mask, torch.finfo(torch.float32).min), dim=-1); del s
    t2 = L.BqTensor3(*p.shape, *p.stride()); y2 = torch.empty_like(p)
    for name, f in [("bfp6", fmt("block_fp", 6, 8, 127)), ("bl8", fmt("block_log", 8, 0, 0, 8))]:
        n = lib.bq_quantize_workspace_bytes(ctypes.byref(f), ctypes.byref(t2)); ws = torch.empty(max(n, 256), dtype=torch.uint8, device=dev)
        ms = timeit(lambda: lib.bq_quantize(ctypes.byref(f), ctypes.byref(t2), p.data_ptr(), y2.data_ptr(), 0, 0, ws.data_ptr(), ws.numel(), L.stream_ptr()))
        out[name + "_probs_f32_GBs"] = round(p.numel() * 8 / ms / 1e6, 1)
    del x, y, ybf, p, y2
if which in ("all", "gemm"):
    for (b, M, N, K) in [(1, 4096, 4096, 4096), (1, 8192, 8192, 8192), (1, 16384, 2048, 2048), (1, 16384, 6144, 2048), (1, 16384, 8192, 2048), (1, 16384, 2048, 8192),
                         (256, 2048, 2048, 64), (256, 2048, 64, 2048)]:
        A = torch.randn(b, M, K, device=dev).to(torch.bfloat16); B = torch.randn(b if b > 1 else 1, N, K, device=dev).to(torch.bfloat16)
        C = torch.empty(b, M, N, device=dev)
        ms = timeit(lambda: lib.bq_gemm_bf16_tn(A.data_ptr(), B.data_ptr(), C.data_ptr(), None, b, M, N, K, K, K, N, M * K, N * K if b > 1 else 0, M * N, L.stream_ptr()), n=10)
        out[f"gemm_{b}x{M}x{N}x{K}_TFLOPs"] = round(2 * b * M * N * K / ms / 1e9, 1)
        out[f"gemm_{b}x{M}x{N}x{K}_ms"] = round(ms, 4)
        if b == 1:
            ms = timeit(lambda: torch.matmul(A[0], B[0].t()), n=10)
            out[f"cublas_{M}x{N}x{K}_TFLOPs"] = round(2 * M * N * K / ms / 1e9, 1)
        del A, B, C
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(f"gpurun_out/bench_kernels_{which}.json", "w"), indent=1)
print(json.dumps(out, indent=1))
if which in ("all", "fused"):
    from llm_mixed_q_b200.models.quantize.quantized_functions.fused_glue import norm_quantize
    from llm_mixed_q_b200.models.quantize.quantized_functions.attention import fused_causal_attention_q
    out = {}
    f6 = ("block_fp", dict(width=6, exponent_width=8, exponent_bias=127))
    for rows, H in [(16384, 2048), (16384, 4096), (16384, 768)]:
        x = torch.randn(rows, H, device=dev); w = torch.ones(H, device=dev); b = torch.zeros(H, device=dev)
        ms = timeit(lambda: norm_quantize(x, w, b, 1e-5, [f6, f6, f6]))
        out[f"ln_quant_{rows}x{H}_us"] = round(ms * 1e3, 1); out[f"ln_quant_{rows}x{H}_GBs"] = round(rows * H * 6 / ms / 1e6, 1)
        ms = timeit(lambda: norm_quantize(x, w, None, 1e-5, [f6]))
        out[f"rms_quant_{rows}x{H}_GBs"] = round(rows * H * 6 / ms / 1e6, 1)
    cfg = {"name": "block_fp", "bypass": False, "is_ptq": True}
    for p in ("data_in", "weight", "bias"):
        cfg.update({f"{p}_width": 6, f"{p}_exponent_width": 8, f"{p}_exponent_bias": 127, f"{p}_block_size": [16] if p == "bias" else [1, 16]})
    for (B, heads, S, d) in [(8, 32, 2048, 64), (2, 32, 2048, 128)]:
        Hh = heads * d
        q = torch.randn(B, S, Hh, device=dev).to(torch.bfloat16); k = torch.randn(B, S, Hh, device=dev).to(torch.bfloat16); v = torch.randn(B, S, Hh, device=dev).to(torch.bfloat16)
        ms = timeit(lambda: fused_causal_attention_q(q, k, v, cfg, heads, B, S, 1.0, out_cfg=cfg), n=10)
        out[f"attention_B{B}h{heads}S{S}d{d}_ms"] = round(ms, 4)
        out[f"attention_B{B}h{heads}S{S}d{d}_Gscores_per_s"] = round(B * heads * S * (S + 1) / 2 / ms / 1e6, 1)
    # GEMM with fused epilogues at the OPT-1.3B layer shapes
    import ctypes as C
    from llm_mixed_q_b200.models.quantize.quantizers.utils import make_format
    fq = make_format("block_fp", width=6, exponent_width=8, exponent_bias=127, b0=1, b1=16)
    for (M, N, K, mode) in [(16384, 2048, 2048, "plain"), (16384, 2048, 2048, "q_n"), (16384, 2048, 2048, "q_m"), (16384, 2048, 2048, "residual"),
                            (16384, 8192, 2048, "relu_q_n"), (16384, 2048, 8192, "residual")]:
        A = torch.randn(M, K, device=dev).to(torch.bfloat16); Bw = (torch.randn(N, K, device=dev) * 0.02).to(torch.bfloat16)
        bias = torch.randn(N, device=dev) * 0.02; res = torch.randn(M, N, device=dev)
        ep = L.BqGemmEpilogue(); ep.bias = bias.data_ptr(); ep.scale = 1.0
        bf = mode not in ("plain", "residual")
        if mode == "residual": ep.residual, ep.ldr = res.data_ptr(), N
        if "q_" in mode: ep.qfmt = C.pointer(fq); ep.qdir = 1 if mode == "q_m" else 0
        if "relu" in mode: ep.act = 1
        ep.out_dtype = 1 if bf else 0
        Cc = torch.empty(M, N, device=dev, dtype=torch.bfloat16 if bf else torch.float32)
        ms = timeit(lambda: lib.bq_gemm_bf16_tn_ex(A.data_ptr(), Bw.data_ptr(), Cc.data_ptr(), C.byref(ep), M, N, K, K, K, N, L.stream_ptr()), n=10)
        out[f"gemm_epi_{mode}_{M}x{N}x{K}_TFLOPs"] = round(2 * M * N * K / ms / 1e9, 1)
        del A, Bw, Cc, res
    json.dump(out, open("gpurun_out/bench_kernels_fused.json", "w"), indent=1)
    print(json.dumps(out, indent=1))
