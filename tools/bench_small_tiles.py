import sys, os, json, ctypes as C, torch
sys.path.insert(0, os.getcwd())
from llm_mixed_q_b200 import _lib as L
lib = L.load(); dev = torch.device("cuda:0")
def timeit(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
out = {}
for (M, N, K) in [(4096, 512, 4096), (4096, 512, 16384), (4096, 2048, 4096), (4096, 1024, 4096), (2048, 768, 768), (512, 512, 512), (16384, 2048, 2048), (16384, 512, 4096)]:
    A = torch.randn(M, K, device=dev).to(torch.bfloat16); B = torch.randn(N, K, device=dev).to(torch.bfloat16); Cc = torch.empty(M, N, device=dev)
    r = {}
    for st in (0, 1):
        lib.bq_set_small_tiles(st)
        ms = timeit(lambda: lib.bq_gemm_bf16_tn(A.data_ptr(), B.data_ptr(), Cc.data_ptr(), None, 1, M, N, K, K, K, N, 0, 0, 0, L.stream_ptr()))
        r["small_tiles" if st else "largest_tile"] = round(ms * 1e3, 1)
        ref = Cc.clone() if st == 0 else ref
    r["bit_identical"] = bool(torch.equal(ref, Cc))
    out[f"{M}x{N}x{K}_us"] = r
print(json.dumps(out, indent=1))
json.dump(out, open("gpurun_out/r02_small_tile_ab.json", "w"), indent=1)
