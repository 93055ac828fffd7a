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


# ------------------------------------------------------------------------------------------------------------------------------
# tensor-parallel decoder layer (BASELINE configs[4]): every Linear column-parallel, exchanges fused into the producing kernels
# ------------------------------------------------------------------------------------------------------------------------------
def _tiny_layer(dev, widths=(6, 6)):
    from llm_mixed_q_b200.models.opt_quantized import OPTQuantizedConfig
    from llm_mixed_q_b200.models.opt_quantized.modeling_opt import OPTQuantizedDecoderLayer

    d = dict(CFG)
    d["data_in_width"], d["weight_width"] = widths
    cfg = OPTQuantizedConfig(hidden_size=256, num_hidden_layers=1, ffn_dim=512, num_attention_heads=4, vocab_size=128,
                             max_position_embeddings=256, quant_config={"default": d})
    torch.manual_seed(21)
    with torch.device(dev):
        layer = OPTQuantizedDecoderLayer(cfg, 0).eval()
        for lin in (layer.self_attn.q_proj, layer.self_attn.k_proj, layer.self_attn.v_proj, layer.self_attn.out_proj, layer.fc1, layer.fc2):
            lin.weight.data.normal_(0, 0.05)
            lin.bias.data.normal_(0, 0.02)
    return layer


def _tp_worker(rank, world, port, backend, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    ndev = torch.cuda.device_count()
    dev = torch.device("cuda", rank % ndev)
    torch.cuda.set_device(dev)
    dist.init_process_group(backend, rank=rank, world_size=world, **({"device_id": dev} if backend == "nccl" else {}))
    try:
        from llm_mixed_q_b200.dist import PeerArena, TensorParallelOPTLayer

        layer = _tiny_layer(dev)                                   # same seed on every rank: identical full layer
        B, S, H = 2, 128, 256
        arena = PeerArena(B * S * 512 * 4, dev, slots=6)
        tp = TensorParallelOPTLayer(layer, arena=arena)            # shards BEFORE any PTQ overwrite
        ok = True
        with torch.no_grad():
            for it in range(3):                                    # 12 takes over 6 slots: exercises slot reuse
                h = torch.randn(B, S, H, device=dev, generator=torch.Generator(device=dev).manual_seed(300 + it))
                ref = layer._fused_forward(h, layer._fused_plan(S))
                out = tp(h, mode="fused").clone()
                torch.cuda.synchronize()
                ok = ok and bool(torch.equal(ref.view(torch.int32), out.view(torch.int32)))
                if backend == "nccl":
                    out2 = tp(h, mode="nccl")
                    ok = ok and bool(torch.equal(ref.view(torch.int32), out2.view(torch.int32)))
        ok = ok and not arena.timed_out()
        ret[rank] = ok
        dist.barrier()
        arena.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_tensor_parallel_layer_two_ranks_bit_identical_shared_device():
    """2 ranks: head-parallel attention + column-parallel Linears with bf16 / fp32 peer stores from the producing kernels; the layer
    output on every rank is bit-identical to the single-GPU fused layer.  (Both ranks share cuda:0 on a one-GPU box.)"""
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_tp_worker, args=(world, _free_port(), "gloo", ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


@pytest.mark.timeout(300)
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (run with gpurun --gpus 2)")
def test_tensor_parallel_layer_two_ranks_distinct_devices_nccl():
    """Same on two REAL devices over NVLink, one rank per GPU, NCCL rendezvous; also checks the NCCL-gather baseline schedule."""
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_tp_worker, args=(world, _free_port(), "nccl", ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


@pytest.mark.timeout(300)
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (run with gpurun --gpus 2)")
def test_fused_column_parallel_two_ranks_distinct_devices():
    """test_fused_column_parallel_two_ranks_bit_identical with one rank per GPU (the worker places rank r on cuda:r when it exists)."""
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), 512, 1024, 384, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_peer_push_copies_strided_slab():
    from llm_mixed_q_b200 import _lib as L

    lib = L.load()
    src = torch.arange(64 * 96, dtype=torch.float32, device="cuda:0").view(64, 96)
    dsts = [torch.full((64, 96), -1.0, device="cuda:0") for _ in range(3)]
    arr = (ctypes.c_void_p * 3)(*[d.data_ptr() + 32 * 4 for d in dsts])
    L.check(lib.bq_peer_push(src.data_ptr() + 32 * 4, arr, 3, 64, 32 * 4, 96 * 4, 96 * 4, L.stream_ptr(src.device)), "push")
    torch.cuda.synchronize()
    for d in dsts:
        assert torch.equal(d[:, 32:64], src[:, 32:64]) and bool((d[:, :32] == -1).all()) and bool((d[:, 64:] == -1).all())
    assert lib.bq_peer_push(src.data_ptr() + 4, arr, 3, 64, 32 * 4, 96 * 4, 96 * 4, None) != 0      # misaligned source
