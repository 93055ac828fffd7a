"""A/B of the operator-API quantized Linear (LinearBlockFP.forward on an fp32 activation), W6A6 and W4A4 block_fp:
  two_launch   quantize kernel (fp32 -> bf16) + tcgen05 GEMM against the bf16 weight cache      (default, bq_linear)
  fused        x-quantizer in the GEMM prologue, one launch                                      (bq_linear_fused)
  packed       quantize kernel + GEMM that decodes (w + 0.5)-bit packed weights in its mainloop  (bq_gemm_packed_tn)
Inputs rotate over > 126 MB of buffers so that nothing is served from L2 by accident.  One JSON object -> gpurun_out/bench_xform.json."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llm_mixed_q_b200 import _lib as L
from llm_mixed_q_b200.models.quantize import get_quantized_cls
from llm_mixed_q_b200.models.quantize.quantized_modules import linear as lin_mod

L.load(); dev = torch.device("cuda:0")
peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}


def cfg(w):
    c = {"name": "block_fp", "bypass": False, "is_ptq": True}
    for p in ("data_in", "weight", "bias"):
        c.update({f"{p}_width": w, f"{p}_exponent_width": 8, f"{p}_exponent_bias": 127, f"{p}_block_size": [16] if p == "bias" else [1, 16]})
    return c


def timeit(fn, n):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n): fn(i)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


out = {}
shapes = [(4096, 4096, 4096), (16384, 2048, 2048), (16384, 2048, 8192), (64, 4096, 4096), (64, 4096, 16384), (16, 4096, 16384), (256, 4096, 16384)]
for w in (6, 4):
    for (M, K, N) in shapes:
        c = cfg(w)
        torch.manual_seed(0)
        lin = get_quantized_cls("linear", c)(K, N, bias=True, config=c).to(dev).eval()
        with torch.no_grad():
            lin.weight.normal_(0, 0.02); lin.bias.normal_(0, 0.02)
        nbuf = max(2, min(8, int(200e6 // (M * K * 4)) + 1))
        xs = [torch.randn(M, K, device=dev) for _ in range(nbuf)]
        key = f"w{w}a{w}_M{M}_K{K}_N{N}"
        flops = 2.0 * M * N * K
        res = {}
        with torch.no_grad():
            ref = lin(xs[0])
            for mode in ("two_launch", "fused", "packed"):
                lin_mod.FUSED_PROLOGUE, lin_mod.PACKED_WEIGHTS = mode == "fused", mode == "packed"
                try:
                    y = lin(xs[0])
                    same = bool(torch.equal(y.view(torch.int32), ref.view(torch.int32)))
                    ms = timeit(lambda i: lin(xs[i % nbuf]), 20 if M >= 4096 else 50)
                    res[mode] = {"ms": round(ms, 5), "TFLOPs": round(flops / ms / 1e9, 1), "bit_identical_to_two_launch": same}
                    if mode == "packed":
                        # the packed form cannot hold the reference's pass-through weights (|w| <= 1e-8 stay UNQUANTISED fp32 values,
                        # block_fp.py:93-94): bq_pack_weight counts them; with N(0, 0.02) weights that is ~4e-7 of the elements
                        res[mode]["pass_through_weight_elements"] = int(lin._wq_packed[2]) if lin._wq_packed else None
                        res[mode]["max_abs_diff_vs_two_launch"] = float((y - ref).abs().max())
                finally:
                    lin_mod.FUSED_PROLOGUE = lin_mod.PACKED_WEIGHTS = False
            bits = lin.packed_bits_per_element()
        res["weight_bytes"] = {"fp32_param": N * K * 4, "bf16_cache": N * K * 2, "packed": int(N * K * bits / 8), "packed_bits_per_element": bits}
        if M <= 256:       # weight-streaming regime: report the weight stream rate of each variant's GEMM-side bytes
            res["two_launch"]["weight_stream_GBs"] = round(N * K * 2 / res["two_launch"]["ms"] / 1e6, 1)
            res["packed"]["weight_stream_GBs"] = round(N * K * bits / 8 / res["packed"]["ms"] / 1e6, 1)
        out[key] = res
        del lin, xs
        torch.cuda.empty_cache()
out["peaks"] = {"hbm_GBs": peaks["hbm_gbs"], "bf16_TFLOPs_burst": peaks["bf16_tflops"]}
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/bench_xform.json", "w"), indent=1)
print(json.dumps(out, indent=1))

# ---- weight-streaming regime, GEMM kernels only (C ABI called directly, pre-quantised bf16 activation, weight buffers rotated so that
# ---- every launch streams its weights from HBM: 3 copies x 54..134 MB > 126 MB L2)
import ctypes
from llm_mixed_q_b200.models.quantize.quantized_modules.linear import pack_weight
from llm_mixed_q_b200.models.quantize.quantizers import block_fp_quantizer
from llm_mixed_q_b200.models.quantize.quantizers.utils import make_format
lib = L.load()
stream = {}
for w in (6, 4, 3):
    for (M, K, N) in [(16, 4096, 16384), (64, 4096, 16384), (128, 4096, 16384), (64, 16384, 4096)]:
        copies = 4
        wqs = [block_fp_quantizer(torch.randn(N, K, device=dev) * 0.02, w, 8, 127, [1, 16], False) for _ in range(copies)]
        packed = [pack_weight(q, w, 8, 127)[0] for q in wqs]
        bf16 = [q.to(torch.bfloat16) for q in wqs]
        del wqs
        xq = block_fp_quantizer(torch.randn(M, K, device=dev), 6, 8, 127, [1, 16], True).to(torch.bfloat16)
        y = torch.empty(M, N, device=dev)
        fmt = make_format("block_fp", width=w, exponent_width=8, exponent_bias=127, b0=1, b1=16)
        sp = L.stream_ptr(dev)
        f_p = lambda i: lib.bq_gemm_packed_tn(xq.data_ptr(), packed[i % copies].data_ptr(), ctypes.byref(fmt), y.data_ptr(), None, M, N, K, K, N, sp)
        f_b = lambda i: lib.bq_gemm_bf16_tn(xq.data_ptr(), bf16[i % copies].data_ptr(), y.data_ptr(), None, 1, M, N, K, K, K, N, 0, 0, 0, sp)
        f_b(0); yb = y.clone(); f_p(0); yp = y.clone()
        ms_p, ms_b = timeit(f_p, 40), timeit(f_b, 40)
        pb, bb = packed[0].numel(), N * K * 2
        stream[f"w{w}_M{M}_K{K}_N{N}"] = {
            "packed_ms": round(ms_p, 5), "bf16_ms": round(ms_b, 5), "speedup_packed_over_bf16": round(ms_b / ms_p, 3),
            "packed_weight_GBs": round(pb / ms_p / 1e6, 1), "bf16_weight_GBs": round(bb / ms_b / 1e6, 1),
            "packed_frac_of_hbm": round(pb / ms_p / 1e6 / peaks["hbm_gbs"], 3), "bf16_frac_of_hbm": round(bb / ms_b / 1e6 / peaks["hbm_gbs"], 3),
            "packed_bytes": pb, "bf16_bytes": bb, "max_abs_diff": float((yb - yp).abs().max())}
        del packed, bf16
        torch.cuda.empty_cache()
out["weight_streaming_gemm_only"] = stream
json.dump(out, open("gpurun_out/bench_xform.json", "w"), indent=1)
print(json.dumps(stream, indent=1))
