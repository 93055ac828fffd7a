#!/usr/bin/env python
"""
Secondary measurements for BASELINE.json configs[3] and configs[4] (bench.py stays the headline, configs[2]).

  python tools/bench_configs.py --config 4 [--format block_minifloat|block_log|both] [--batch 2]
      Llama-7B shape (32 layers, H 4096, I 11008, h 32, d 128, vocab 32000), random init, W4A4 block_minifloat /
      block_log (configs/llama_w4a4_*.toml), seq 2048, full forward incl. fp32 lm_head + shifted CE loss.
      Data-parallel: every rank runs its own replica (weak scaling, no data-path collective).  tokens/s.
  python tools/bench_configs.py --config 5 [--tokens 4096]
      OPT-6.7B layer GEMMs (H 4096, F 16384) under the per-layer mixed-precision block_fp config
      (configs/opt_6.7b_mixed_bfp.toml, section-4.4 search format), column-parallel over the ranks of the job:
      quantize x -> tcgen05 GEMM on this rank's N/g rows -> all-gather of the fp32 column slabs.
      Checks the gathered result is BIT-IDENTICAL to the one-GPU module, then times it (CUDA events, max over ranks).

Launch N>1 as:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_configs.py ...
Each run prints one JSON line per measurement on rank 0 and appends it to gpurun_out/bench_configs.jsonl.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

SEQ = 2048


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


def emit(line, rank):
    if rank != 0:
        return
    print(json.dumps(line), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "bench_configs.jsonl"), "a") as f:
        f.write(json.dumps(line) + "\n")


def setup():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    return world, rank, dev


def max_over_ranks(ms, world, dev):
    if world == 1:
        return ms
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def barrier(world):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


# ---------------------------------------------------------------------------------------------------
def config4(args, world, rank, dev):
    from llm_mixed_q_b200 import _lib as L
    from llm_mixed_q_b200.models.llama_quantized import LlamaQuantizedConfig, LlamaQuantizedForCausalLM

    kinds = ["block_minifloat", "block_log"] if args.format == "both" else [args.format]
    for kind in kinds:
        toml_path = os.path.join(ROOT, "configs", f"llama_w4a4_{kind}.toml")
        # N(0,0.02) weights all quantise to 0 under block_minifloat (shared bias clamps at >= 0):
        # the timed run uses the x64-scaled init so the arithmetic is not degenerate
        init = 1.28 if kind == "block_minifloat" else 0.02
        cfg = LlamaQuantizedConfig(quant_config=toml_path, initializer_range=init,
                                   num_hidden_layers=args.layers or 32)
        torch.manual_seed(0)
        t0 = time.time()
        with torch.device(dev):
            model = LlamaQuantizedForCausalLM(cfg).eval()
        g = torch.Generator(device="cpu").manual_seed(rank)
        ids = torch.randint(0, cfg.vocab_size, (args.batch, SEQ), generator=g).to(dev)
        K, W = args.steps, max(args.warmup, 2)
        with torch.no_grad():
            for _ in range(W):
                out = model(input_ids=ids, labels=ids)
            barrier(world)
            l0 = sum(L.launch_counts().values())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(K):
                out = model(input_ids=ids, labels=ids)
            e1.record()
            barrier(world)
        ms = max_over_ranks(e0.elapsed_time(e1), world, dev) / K
        launches = sum(L.launch_counts().values()) - l0
        # the same forward replayed from a CUDA graph (llm_mixed_q_b200/utils/graphs.py): ~700 launches per step issued from Python
        # make the eager number follow the host's speed
        graph_ms, graph_err = None, None
        if args.graph:
            from llm_mixed_q_b200.utils.graphs import GraphedForward
            runner = GraphedForward(model, args.batch, SEQ, device=dev)
            graph_err = runner.error
            if runner.graph is not None:
                for _ in range(2):
                    runner(ids)
                barrier(world)
                e0.record()
                for _ in range(K):
                    runner(ids)
                e1.record()
                barrier(world)
                graph_ms = max_over_ranks(e0.elapsed_time(e1), world, dev) / K
                assert abs(float(runner.loss) - float(out.loss)) <= 1e-6 * abs(float(out.loss)), (float(runner.loss), float(out.loss))
            del runner
        eager_ms = ms
        if graph_ms is not None:
            ms = graph_ms
        tokens = args.batch * SEQ * world
        # algorithmic FLOPs per token: 7 Linears + 2 matmuls per layer + lm_head
        H, I, Lyr, V, h, d = cfg.hidden_size, cfg.intermediate_size, cfg.num_hidden_layers, cfg.vocab_size, 32, 128
        flops = 2 * args.batch * SEQ * (Lyr * (4 * H * H + 3 * H * I) + H * V) + Lyr * 2 * 2 * args.batch * h * SEQ * SEQ * d
        _, burst, sus, src = peaks()
        emit({"metric": f"W4A4-{kind} fwd tokens/s (Llama-7B)", "value": tokens / (ms / 1e3), "unit": "tokens/s", "n_gpus": world,
              "steps": K, "warmup": W, "ms_per_step": ms, "scaling": "weak", "dtype": "bf16", "data": "synthetic",
              "config": {"workload": f"Llama-7B shape W4A4 {kind} (block 16) full forward, seq 2048, batch {args.batch} per GPU",
                         "toml": os.path.relpath(toml_path, ROOT), "layers": Lyr, "init_std": init,
                         "parallelism": f"dp{world} (independent replicas)", "loss": float(out.loss)},
              "algorithmic_TFLOPs": flops * world / (ms / 1e3) / 1e12, "frac_of_bf16_sustained": flops / (ms / 1e3) / 1e12 / sus,
              "peak_source": src, "gpu_launches_per_step": launches // K, "build_s": round(time.time() - t0, 1),
              "eager_ms_per_step": eager_ms, "graph_replay_ms_per_step": graph_ms, "graph_error": graph_err}, rank)
        del model, out
        torch.cuda.empty_cache()


# ---------------------------------------------------------------------------------------------------
def config1(args, world, rank, dev):
    """BASELINE configs[0]: OPT-125M shape (the config class defaults), random init, W6A6 block_fp, 1 x 2048 tokens — the case the
    reference itself runs on a CPU (oracle/gen_golden_opt125m.py: 44.7 s on 8 cores = 46 tokens/s) — and the same model at batch 8.
    Parity of this model against the reference's own forward: tests/test_gpu_models.py::test_opt125m_config1_*."""
    from llm_mixed_q_b200.models.opt_quantized import OPTQuantizedConfig, OPTQuantizedForCausalLM
    from llm_mixed_q_b200.utils.graphs import GraphedForward

    cfg = OPTQuantizedConfig(quant_config=os.path.join(ROOT, "configs", "bfp_w6a6.toml"), tie_word_embeddings=False)
    torch.manual_seed(0)
    with torch.device(dev):
        model = OPTQuantizedForCausalLM(cfg).eval()
    K, W = args.steps, max(args.warmup, 2)
    for B in (1, 8):
        g = torch.Generator(device="cpu").manual_seed(rank)
        ids = torch.randint(0, cfg.vocab_size, (B, SEQ), generator=g).to(dev)
        runner = GraphedForward(model, B, SEQ, device=dev)
        for _ in range(W):
            runner(ids)
        barrier(world)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            runner(ids)
        e1.record()
        barrier(world)
        ms = max_over_ranks(e0.elapsed_time(e1), world, dev) / K
        emit({"metric": "W6A6-BFP fwd tokens/s (OPT-125M)", "value": B * SEQ * world / (ms / 1e3), "unit": "tokens/s", "n_gpus": world,
              "steps": K, "warmup": W, "ms_per_step": ms, "scaling": "weak", "dtype": "bf16", "data": "synthetic",
              "config": {"workload": f"OPT-125M shape W6A6 block_fp full forward + loss, seq 2048, batch {B} per GPU", "mode":
                         "cuda-graph replay" if runner.graph is not None else "eager", "loss": float(runner.loss)}}, rank)
        del runner


# ---------------------------------------------------------------------------------------------------
def config_bert(args, world, rank, dev):
    """BERT-base shape (12 layers, H 768, 12 heads x 64, I 3072) W6A6 block_fp sequence classification, seq 512: the bidirectional /
    key-padded fused attention (bq_attention_masked) against the op-by-op attention of this package (S x S scores through HBM).
    One third of the sequences is right-padded to 384 tokens."""
    from llm_mixed_q_b200 import _lib as L
    from llm_mixed_q_b200.models.bert_quantized import BertQuantizedConfig, BertQuantizedForSequenceClassification

    S, B = 512, args.batch
    toml_path = os.path.join(ROOT, "configs", "bfp_w6a6.toml")
    cfg = BertQuantizedConfig(quant_config=toml_path, num_labels=2)
    torch.manual_seed(0)
    with torch.device(dev):
        model = BertQuantizedForSequenceClassification(cfg).eval()
    g = torch.Generator(device="cpu").manual_seed(rank)
    ids = torch.randint(1000, cfg.vocab_size, (B, S), generator=g).to(dev)
    am = torch.ones(B, S, dtype=torch.long, device=dev)
    am[::3, 384:] = 0
    K, W = args.steps, max(args.warmup, 2)
    res = {}
    for fused in (True, False):
        model.bert.fused_attention = fused
        with torch.no_grad():
            for _ in range(W):
                out = model(input_ids=ids, attention_mask=am)
            barrier(world)
            n0 = L.launch_counts()["attention_causal_kernel"]
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(K):
                out = model(input_ids=ids, attention_mask=am)
            e1.record()
            barrier(world)
        res[fused] = (max_over_ranks(e0.elapsed_time(e1), world, dev) / K, (L.launch_counts()["attention_causal_kernel"] - n0) // K,
                      out.logits.float().cpu())
    ms = res[True][0]
    emit({"metric": "W6A6-BFP fwd tokens/s (BERT-base, seq 512)", "value": B * S * world / (ms / 1e3), "unit": "tokens/s", "n_gpus": world,
          "steps": K, "warmup": W, "ms_per_step": ms, "scaling": "weak", "dtype": "bf16", "data": "synthetic",
          "config": {"workload": f"BERT-base shape W6A6 block_fp sequence classification, seq 512, batch {B} per GPU, 1/3 of the rows "
                                 "right-padded to 384 tokens (key-padding mask)", "toml": os.path.relpath(toml_path, ROOT)},
          "fused_attention_launches_per_step": res[True][1], "op_by_op_attention_ms_per_step": res[False][0],
          "op_by_op_tokens_per_s": B * S * world / (res[False][0] / 1e3),
          "max_abs_dlogit_fused_vs_op_by_op": float((res[True][2] - res[False][2]).abs().max())}, rank)


# ---------------------------------------------------------------------------------------------------
def config5(args, world, rank, dev):
    from llm_mixed_q_b200 import _lib as L
    from llm_mixed_q_b200.dist import ColumnParallelLinear, PeerArena
    from llm_mixed_q_b200.models.opt_quantized import parse_opt_quantized_config
    from llm_mixed_q_b200.models.quantize import get_quantized_cls

    qc = parse_opt_quantized_config(os.path.join(ROOT, "configs", "opt_6.7b_mixed_bfp.toml"), 32)
    H, F_ = 4096, 16384
    M = args.tokens
    shapes = [("self_attn.q_proj", H, H), ("self_attn.k_proj", H, H), ("self_attn.v_proj", H, H), ("self_attn.out_proj", H, H),
              ("fc1", H, F_), ("fc2", F_, H)]
    _, burst, sus, src = peaks()
    layer = args.layer
    total_ms, total_flops, total_gemm_ms, total_fused_ms = 0.0, 0, 0.0, 0.0
    arena = None
    if not args.no_peer:
        try:
            arena = PeerArena(M * F_ * 4, dev)          # two slots of the largest gathered output (fc1)
        except Exception as e:                          # e.g. IPC not permitted in this container: NCCL numbers still print
            print(f"[rank {rank}] PeerArena unavailable: {type(e).__name__}: {e}", flush=True)
    for name, K, N in shapes:
        node = qc[f"model_layer_{layer}"]
        for part in name.split("."):
            node = node[part]
        torch.manual_seed(1234)                     # every rank builds the SAME full module and input
        with torch.device(dev):
            full = get_quantized_cls("linear", node)(K, N, bias=True, config=node).eval()
            full.bias.data.normal_(0, 0.02)
        x = torch.randn(M, K, device=dev, generator=torch.Generator(device=dev).manual_seed(7))
        cp = ColumnParallelLinear.from_linear(full)          # shards BEFORE the PTQ overwrite: blocks never cross the cut
        cp_local = ColumnParallelLinear(cp.local, N, gather_output=False)
        cp_fused = ColumnParallelLinear(cp.local, N, arena=arena) if arena is not None else None
        with torch.no_grad():
            y_full = full(x)
            y_cp = cp(x)
            identical = bool(torch.equal(y_full.view(torch.int32), y_cp.view(torch.int32)))
            fused_identical = None
            if cp_fused is not None:
                assert cp_fused._fused_ok(x)
                y_f = cp_fused(x)
                torch.cuda.synchronize()
                fused_identical = bool(torch.equal(y_full.view(torch.int32), y_f.view(torch.int32))) and not arena.timed_out()
                del y_f
        del y_full, y_cp

        def timed(fn, iters=args.steps):
            with torch.no_grad():
                for _ in range(max(args.warmup, 3)):
                    fn()
                barrier(world)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(iters):
                    fn()
                b.record()
                barrier(world)
            return max_over_ranks(a.elapsed_time(b), world, dev) / iters

        ms = timed(lambda: cp(x))
        ms_local = timed(lambda: cp_local(x))
        ms_fused = timed(lambda: cp_fused(x)) if cp_fused is not None else None
        flops = 2 * M * N * K
        total_ms += ms
        total_gemm_ms += ms_local
        total_fused_ms += ms_fused or 0.0
        total_flops += flops
        emit({"metric": "column-parallel q-GEMM TFLOP/s (OPT-6.7B mixed block_fp)", "op": name, "M": M, "K": K, "N": N, "n_gpus": world,
              "x_width": node["data_in_width"], "w_width": node["weight_width"], "ms": ms, "ms_quantize_plus_gemm": ms_local,
              "ms_all_gather": ms - ms_local, "value": flops / (ms / 1e3) / 1e12, "unit": "TFLOP/s (whole job)",
              "frac_of_bf16_sustained_x_gpus": flops / (ms / 1e3) / 1e12 / (sus * world),
              "gemm_only_frac": flops / (ms_local / 1e3) / 1e12 / (sus * world), "bit_identical_to_1gpu": identical,
              "ms_fused_peer_store": ms_fused, "fused_TFLOPs": (flops / (ms_fused / 1e3) / 1e12) if ms_fused else None,
              "fused_bit_identical_to_1gpu": fused_identical, "fused_timed_out": arena.timed_out() if arena is not None else None,
              "all_gather_bytes_per_rank": M * (N // world) * 4 * (world - 1), "peak_source": src}, rank)
        assert identical, f"{name}: column-parallel result differs from the 1-GPU module"
        assert fused_identical in (None, True), f"{name}: fused peer-store result differs from the 1-GPU module"
        del full, cp, cp_local, cp_fused, x
        torch.cuda.empty_cache()
    emit({"metric": "column-parallel q-GEMM TFLOP/s (OPT-6.7B mixed block_fp)", "op": "layer total (6 Linears)", "M": M, "n_gpus": world,
          "layer": layer, "ms": total_ms, "ms_quantize_plus_gemm": total_gemm_ms, "value": total_flops / (total_ms / 1e3) / 1e12,
          "ms_fused_peer_store": total_fused_ms or None,
          "fused_TFLOPs": (total_flops / (total_fused_ms / 1e3) / 1e12) if total_fused_ms else None,
          "unit": "TFLOP/s (whole job)", "frac_of_bf16_sustained_x_gpus": total_flops / (total_ms / 1e3) / 1e12 / (sus * world),
          "scaling": "strong", "gpu_launches": sum(L.launch_counts().values())}, rank)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True, choices=[4, 5, 0, 1], help="1 / 4 / 5: BASELINE configs[0] / [3] / [4]; 0: BERT-base")
    ap.add_argument("--format", default="both", choices=["block_minifloat", "block_log", "both"])
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--layers", type=int, default=None, help="debug: fewer layers")
    ap.add_argument("--tokens", type=int, default=4096)
    ap.add_argument("--layer", type=int, default=0)
    ap.add_argument("--no-peer", action="store_true", help="config 5: skip the fused peer-store path (NCCL all-gather only)")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--graph", action="store_true", help="config 4: also time the forward replayed from a CUDA graph (reported value)")
    ap.add_argument("--warmup", type=int, default=2)
    args = ap.parse_args()
    world, rank, dev = setup()
    from llm_mixed_q_b200 import _lib as L

    L.load()
    {4: config4, 5: config5, 0: config_bert, 1: config1}[args.config](args, world, rank, dev)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
