"""GPU tests of the fused column-parallel Linear: GEMM epilogue stores into peer-mapped buffers + flag barrier
(include/bq.h: bq_gemm_epilogue.replicas, bq_ipc_*, bq_peer_barrier; llm_mixed_q_b200/dist.py: PeerArena).

The box the driver tests on has ONE GPU, so the two ranks of the multi-process test share cuda:0 (CUDA IPC and
system-scope flags work between processes on one device; rendezvous over gloo).  On a multi-GPU box the same test body runs
one rank per GPU (tools/bench_configs.py --config 5 does that under torchrun with NCCL)."""
import ctypes
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

CFG = {"name": "block_fp", "bypass": False, "is_ptq": True}
for _p in ("data_in", "weight", "bias"):
    CFG.update({f"{_p}_width": 6, f"{_p}_exponent_width": 8, f"{_p}_exponent_bias": 127,
                f"{_p}_block_size": [16] if _p == "bias" else [1, 16]})


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _build(dev, K, N):
    from llm_mixed_q_b200.models.quantize import get_quantized_cls

    torch.manual_seed(11)
    with torch.device(dev):
        full = get_quantized_cls("linear", CFG)(K, N, bias=True, config=CFG).eval()
        full.bias.data.normal_(0, 0.02)
    return full


def test_epilogue_replica_stores_single_process():
    """n_replicas > 0 with plain local pointers: every replica receives exactly the bytes of C, fp32 and bf16."""
    from llm_mixed_q_b200.models.quantize.quantized_modules.linear import operand_format, quantize_operand_bf16

    dev = torch.device("cuda:0")
    K, N, M = 256, 320, 200
    full = _build(dev, K, N)
    x = torch.randn(M, K, device=dev, generator=torch.Generator(device=dev).manual_seed(3))
    with torch.no_grad():
        y_ref = full(x)
        kind, kw, bs = operand_format(CFG, "data_in")
        xq = quantize_operand_bf16(x, kind, kw, bs, True)
        wide = torch.full((M, 3 * N), float("nan"), device=dev)
        reps = [torch.full((M, 3 * N), float("nan"), device=dev) for _ in range(3)]
        out = wide[:, N:2 * N]
        full.forward_prequantized(xq, out=out, peer_out_ptrs=[r.data_ptr() + N * 4 for r in reps])
    torch.cuda.synchronize()
    assert torch.equal(out.view(torch.int32), y_ref.view(torch.int32))
    for r in reps:
        assert torch.equal(r[:, N:2 * N].view(torch.int32), y_ref.view(torch.int32))
        assert torch.isnan(r[:, :N]).all() and torch.isnan(r[:, 2 * N:]).all()      # nothing outside the slab is touched
    assert torch.isnan(wide[:, :N]).all() and torch.isnan(wide[:, 2 * N:]).all()


def test_ipc_export_reports_offset_inside_allocation():
    from llm_mixed_q_b200 import _lib as L

    lib = L.load()
    buf = torch.zeros(4 << 20, dtype=torch.uint8, device="cuda:0")
    h0, h1 = L.BqIpcHandle(), L.BqIpcHandle()
    L.check(lib.bq_ipc_export(buf.data_ptr(), ctypes.byref(h0)), "export")
    L.check(lib.bq_ipc_export(buf.data_ptr() + 4096, ctypes.byref(h1)), "export")
    assert bytes(h0.reserved) == bytes(h1.reserved)
    assert h1.offset == h0.offset + 4096 and h0.size >= 4 << 20
    with pytest.raises(ValueError):
        L.check(lib.bq_ipc_export(None, ctypes.byref(h0)), "export")


def test_peer_barrier_world1_and_bad_args():
    from llm_mixed_q_b200 import _lib as L

    lib = L.load()
    flags = torch.zeros(64, dtype=torch.int32, device="cuda:0")
    sig = (ctypes.c_void_p * 1)(flags.data_ptr())
    L.check(lib.bq_peer_barrier(sig, 0, 1, 1, 100, L.stream_ptr(flags.device)), "barrier")
    torch.cuda.synchronize()
    assert int(flags.abs().sum()) == 0
    assert lib.bq_peer_barrier(sig, 1, 1, 1, 100, None) != 0          # rank out of range
    assert lib.bq_peer_barrier(sig, 0, 9, 1, 100, None) != 0          # more than one NVSwitch domain


def _worker(rank, world, port, K, N, M, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    ndev = torch.cuda.device_count()
    dev = torch.device("cuda", rank % ndev)
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from llm_mixed_q_b200 import _lib as L
        from llm_mixed_q_b200.dist import ColumnParallelLinear, PeerArena

        full = _build(dev, K, N)                                  # same seed on every rank: identical full module
        arena = PeerArena(M * N * 4, dev)
        cp = ColumnParallelLinear.from_linear(full, arena=arena)  # shards BEFORE the PTQ overwrite
        ok, n_barriers0 = True, L.launch_counts()["peer_barrier_kernel"]
        outs = []
        with torch.no_grad():
            for it in range(4):                                   # 4 calls over 2 slots: exercises slot reuse
                x = torch.randn(M, K, device=dev, generator=torch.Generator(device=dev).manual_seed(100 + it))
                y_full = full(x)
                y_cp = cp(x)
                outs.append((y_full, y_cp.clone()))
        torch.cuda.synchronize()
        for y_full, y_cp in outs:
            ok = ok and bool(torch.equal(y_full.view(torch.int32), y_cp.view(torch.int32)))
        ok = ok and not arena.timed_out()
        ok = ok and (L.launch_counts()["peer_barrier_kernel"] - n_barriers0 == 4)
        ok = ok and cp._fused_ok(x)
        ret[rank] = ok
        dist.barrier()
        arena.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_fused_column_parallel_two_ranks_bit_identical():
    """2 ranks: each computes half of the columns and stores them into BOTH ranks' gathered buffers from the GEMM epilogue;
    after the flag barrier every rank holds a result bit-identical to the single-GPU module."""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), 512, 1024, 384, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}
