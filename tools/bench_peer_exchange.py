"""How fast can one rank deliver a column slab to every peer?  Three ways of the same exchange (each rank sends its [M, N/g] bf16 slab
of a gathered [M, N] buffer to all g - 1 peers), timed with CUDA events, max over ranks:
  push     bq_peer_push: SM copy kernel, 16-byte packets, stores straight into the peer-mapped buffers (what TensorParallelOPTLayer uses
           for the attention slab; the GEMM epilogues store the same way)
  ce       cudaMemcpy2DAsync per peer on its own stream (copy engines over NVLink)
  nccl     all_gather_into_tensor of the contiguous slabs (+ nothing else)
each followed by the barrier that makes the data visible (peer flag barrier / stream joins / NCCL's own).
Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_peer_exchange.py"""
import ctypes, json, os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llm_mixed_q_b200 import _lib as L
from llm_mixed_q_b200.dist import PeerArena

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
lib = L.load()
rt = ctypes.CDLL("libcudart.so.12")
rt.cudaMemcpy2DAsync.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
rt.cudaMemcpy2DAsync.restype = ctypes.c_int

def mx(v):
    t = torch.tensor([v], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t[0])

def timed(fn, iters=10):
    for _ in range(3): fn()
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); dist.barrier(); torch.cuda.synchronize()
    return mx(a.elapsed_time(b) / iters)

out = {"world": world}
streams = [torch.cuda.Stream(dev) for _ in range(world)]
for (M, N, esz, name) in [(4096, 16384, 2, "fc1_bf16_4096"), (4096, 4096, 4, "h_fp32_4096"), (4096, 4096, 2, "attn_bf16_4096"), (16384, 16384, 2, "fc1_bf16_16384")]:
    Nl = N // world
    arena = PeerArena(M * N * esz, dev, slots=2)
    dt = torch.bfloat16 if esz == 2 else torch.float32
    full, bases = arena.take((M, N), dt)
    slab = full[:, rank * Nl:(rank + 1) * Nl]
    slab.normal_()
    sent = M * Nl * esz * (world - 1)
    def push():
        arena.push(slab, bases, rank * Nl * esz); arena.barrier()
    def ce():
        cur = torch.cuda.current_stream(dev)
        ev = torch.cuda.Event(); ev.record(cur)
        k = 0
        for r, b in enumerate(bases):
            if r == rank: continue
            st = streams[k]; k += 1
            st.wait_event(ev)
            rc = rt.cudaMemcpy2DAsync(b + rank * Nl * esz, N * esz, slab.data_ptr(), N * esz, Nl * esz, M, 3, st.cuda_stream)
            assert rc == 0, rc
            e2 = torch.cuda.Event(); e2.record(st); cur.wait_event(e2)
        arena.barrier()
    loc = slab.contiguous(); gathered = torch.empty(world, M, Nl, dtype=dt, device=dev)
    def nccl():
        dist.all_gather_into_tensor(gathered.view(-1), loc.view(-1))
    r = {}
    for nm, fn in (("push", push), ("ce", ce), ("nccl", nccl)):
        ms = timed(fn)
        r[nm + "_ms"] = round(ms, 4); r[nm + "_GBs_sent_per_rank"] = round(sent / ms / 1e6, 1)
    r["MB_sent_per_rank"] = round(sent / 1e6, 1)
    out[name] = r
    arena.close()
if rank == 0:
    print(json.dumps(out, indent=1))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/r02_peer_exchange_n{world}.json", "w"), indent=1)
dist.destroy_process_group()
