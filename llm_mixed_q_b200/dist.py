"""
Multi-GPU plumbing for the two places the hot path shards.  One process per GPU,
`torch.distributed` (NCCL on GPUs; gloo in the CPU tests).  The reference has no counterpart: it runs the
emulation on one device, or lets `accelerate` place layers (`cli/eval_lm.py`), which is not on the hot path.

1. Data-parallel forward (configs 3, 4): sequences are independent and blocks never cross the batch dim
   (reference quantizers/utils.py:220-222), so every rank runs a full replica on its own batches and the
   only exchange is ONE all-reduce of (sum of loss*tokens, number of samples) for the perplexity
   (reduction of reference eval/eval_lm.py:41-63).  `shard_batches` / `dp_perplexity`.

2. Column-parallel quantized Linear (config 5): `W[N, K]` is blocked [1,16] along K and the bias [16] along
   N (reference quantized_modules/linear.py:113-143), so a split along N at multiples of 16 keeps every block
   inside one shard.  Rank r holds rows [r*N/g, (r+1)*N/g) of the weight (quantised locally — identical
   values to quantising the full matrix, because no block crosses the cut; block_log, whose all-zero blocks take a
   tensor-global minimum, is quantised on the full matrix before it is cut), computes its column slab
   of y from the replicated x, and the slabs are all-gathered.  The K-reduction of every output element is
   done by the same kernel in the same order as on one GPU, so the result is bit-identical to 1 GPU.
   `ColumnParallelLinear`.  Two exchange implementations behind the same module:
     * `all_gather_columns` — NCCL all-gather of the slabs + one permute (the library baseline), and
     * `PeerArena` (fused, the default on NVLink-connected GPUs) — the GEMM epilogue stores every output tile into all
       ranks' gathered [M, N] buffers while the remaining tiles are still being multiplied (bq_gemm_bf16_tn_ex
       `replicas`, include/bq.h), followed by one flag barrier over peer memory (bq_peer_barrier).  No NCCL call, no
       staging slab, no permute; the bytes that cross NVLink are the all-gather's own (M*N/g*4*(g-1) per rank).
"""
from __future__ import annotations

import ctypes
import math
from typing import Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn


# --------------------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------------------
def _world(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def shard_range(n: int, world: int, rank: int, align: int = 16) -> Tuple[int, int]:
    """Rows [lo, hi) of an N-row weight owned by `rank`.  N/world must be a multiple of `align` (= the bias
    block, so that bias blocks and weight rows split at the same place)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    if n % world != 0 or (n // world) % align != 0:
        raise ValueError(f"out_features={n} cannot be split over {world} ranks in multiples of {align}: "
                         "a bias block would straddle two shards and change its shared exponent")
    per = n // world
    return rank * per, (rank + 1) * per


def shard_batches(num_batches: int, world: int, rank: int) -> range:
    """Round-robin assignment of batch indices to ranks (batch i -> rank i % world)."""
    return range(rank, num_batches, world)


# --------------------------------------------------------------------------------------------------
# 1. data-parallel perplexity
# --------------------------------------------------------------------------------------------------
def perplexity_from_sums(loss_sum: float, num_samples: int, seq_len: int) -> dict:
    """Final reduction of reference eval/eval_lm.py:57-71: exp(sum_i(loss_i * B * S) / (S * N))."""
    reduced = loss_sum / (seq_len * num_samples) if num_samples else float("nan")
    try:
        ppl = math.exp(reduced)
    except OverflowError:
        ppl = float("inf")
    return {"loss": reduced, "perplexity": ppl, "num_samples": num_samples, "seq_len": seq_len}


@torch.no_grad()
def dp_perplexity(model, batches: Sequence[dict], group=None, input_device=None) -> dict:
    """
    Perplexity of `model` over `batches` (each {"input_ids": [B,S], "labels": [B,S], ...}, constant B and S),
    data-parallel: rank r evaluates batches r, r+g, r+2g, ... and one all-reduce combines the two sums.
    With one process this is exactly the loop of reference eval/eval_lm.py:41-63.
    """
    world, rank = _world(group)
    model.eval()
    dev = input_device if input_device is not None else next(model.parameters()).device
    loss_sum, n_samples, seq_len, batch_size = 0.0, 0, None, None
    for i in shard_batches(len(batches), world, rank):
        batch = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in batches[i].items()}
        b, s = batch["input_ids"].shape
        if seq_len is None:
            seq_len, batch_size = s, b
        assert s == seq_len, f"sequence length is not a constant current seq_len = {s} != {seq_len}"
        out = model(**batch)
        loss = out.loss if hasattr(out, "loss") else out[0]
        loss_sum += float(loss) * b * s
        n_samples += b
    if world > 1:
        red_dev = dev if (dist.get_backend(group) == "nccl") else torch.device("cpu")
        t = torch.tensor([loss_sum, float(n_samples), float(seq_len or 0)], dtype=torch.float64, device=red_dev)
        mx = t[2:].clone()
        dist.all_reduce(t[:2], op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
        loss_sum, n_samples, seq_len = float(t[0]), int(t[1]), int(mx[0])
    res = perplexity_from_sums(loss_sum, n_samples, seq_len or 0)
    res["batch_size"] = batch_size
    return res


# --------------------------------------------------------------------------------------------------
# 2. column-parallel quantized Linear
# --------------------------------------------------------------------------------------------------
def all_gather_columns(y_local: torch.Tensor, group=None) -> torch.Tensor:
    """[..., N/g] slabs on g ranks -> [..., N] on every rank (slab r occupies columns [r*N/g, (r+1)*N/g))."""
    world, _ = _world(group)
    if world == 1:
        return y_local
    lead, nl = y_local.shape[:-1], y_local.shape[-1]
    y2 = y_local.reshape(-1, nl).contiguous()
    m = y2.shape[0]
    buf = torch.empty((world * m, nl), dtype=y2.dtype, device=y2.device)     # rank-major: rows [r*m, (r+1)*m) = slab r
    dist.all_gather_into_tensor(buf, y2, group=group)
    return buf.view(world, m, nl).permute(1, 0, 2).reshape(*lead, world * nl)


# peer allocations mapped into this process: (peer rank, 64-byte IPC handle) -> [base address, users].  CUDA maps an
# allocation once per process; two arenas that torch carved out of the same cudaMalloc block share the mapping.
_IPC_MAPPINGS: dict = {}


def _ipc_import(lib, L, peer_rank: int, raw: bytes):
    h = L.BqIpcHandle.from_buffer_copy(raw)
    key = (peer_rank, bytes(h.reserved))
    ent = _IPC_MAPPINGS.get(key)
    if ent is None:
        base, ptr = ctypes.c_void_p(), ctypes.c_void_p()
        L.check(lib.bq_ipc_import(ctypes.byref(h), ctypes.byref(base), ctypes.byref(ptr)), "bq_ipc_import")
        ent = _IPC_MAPPINGS[key] = [base.value, 0]
    ent[1] += 1
    return key, ent[0] + h.offset


def _ipc_release(lib, key):
    ent = _IPC_MAPPINGS.get(key)
    if ent is None:
        return
    ent[1] -= 1
    if ent[1] <= 0:
        lib.bq_ipc_release(ent[0])
        del _IPC_MAPPINGS[key]


class PeerArena:
    """
    Symmetric, peer-mapped output memory for the fused all-gather: every rank owns `slots` buffers of `slot_bytes`
    plus one flag block, allocated by torch (one fresh allocation) and mapped into every other rank with CUDA IPC
    (bq_ipc_export / bq_ipc_import).  Handles travel through `torch.distributed` (all_gather_object), never the data.

    Protocol of one use (`ColumnParallelLinear._forward_fused`): take the next slot (round-robin over >= 2 slots), let the
    GEMM epilogue write this rank's column slab into the slot of EVERY rank, then `barrier()`.  After the barrier the
    local slot holds the complete [M, N] result.  Reusing a slot is safe without a second barrier: a rank starts writing
    slot s of call k only after it passed the barrier of call k-1, which every peer enters only after (stream order) the
    consumers of its call k-2 result — the previous user of slot s — were enqueued ahead of it.
    """

    TIMEOUT_MS = 20000

    def __init__(self, slot_bytes: int, device, group=None, slots: int = 2):
        from . import _lib as L

        if slots < 2:
            raise ValueError("PeerArena needs at least two slots (see the reuse argument in the class docstring)")
        self.group = group
        self.world, self.rank = _world(group)
        if self.world > 8:
            raise ValueError("PeerArena spans one NVSwitch domain (<= 8 GPUs)")
        self.device = torch.device(device)
        self.slot_bytes = (int(slot_bytes) + 255) // 256 * 256
        self.slots = slots
        self._flag_bytes = 1024
        self.lib = L.load()
        self._L = L
        # one private allocation: flags first, then the slots (>= 1 MB so torch gives it its own cudaMalloc block)
        total = self._flag_bytes + self.slots * self.slot_bytes
        self.buf = torch.zeros(max(total, 2 << 20), dtype=torch.uint8, device=self.device)
        torch.cuda.synchronize(self.device)
        h = L.BqIpcHandle()
        L.check(self.lib.bq_ipc_export(self.buf.data_ptr(), ctypes.byref(h)), "bq_ipc_export")
        mine = bytes(h)
        gathered = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(gathered, mine, group=group)
        else:
            gathered[0] = mine
        self._bases = []
        self.ptrs = []
        for r, raw in enumerate(gathered):
            if r == self.rank:
                self.ptrs.append(self.buf.data_ptr())
                continue
            key, ptr = _ipc_import(self.lib, L, r, raw)
            self._bases.append(key)
            self.ptrs.append(ptr)
        self._signals = (ctypes.c_void_p * self.world)(*self.ptrs)
        self.epoch = 0
        self._next = 0
        # sticky error word in mapped pinned host memory: a barrier that gives up waiting writes 1 here (bq_peer_barrier_ex), and the
        # next take() / barrier() on the host raises instead of handing out a partially filled result
        self._host_flag = torch.zeros(1, dtype=torch.int32).pin_memory()
        # nobody signals into a flag block before its owner zeroed it
        if self.world > 1:
            dist.barrier(group=group)

    def take(self, shape, dtype=torch.float32):
        """Next slot as a local tensor of `shape`, and the address of the same slot on every rank."""
        self._raise_if_timed_out()
        n = 1
        for d in shape:
            n *= int(d)
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        if nbytes > self.slot_bytes:
            raise ValueError(f"result of {nbytes} bytes does not fit a PeerArena slot of {self.slot_bytes} bytes")
        off = self._flag_bytes + self._next * self.slot_bytes
        self._next = (self._next + 1) % self.slots
        local = self.buf[off:off + nbytes].view(dtype).view(*shape)
        return local, [p + off for p in self.ptrs]

    def _raise_if_timed_out(self):
        if int(self._host_flag[0]) != 0:
            raise RuntimeError(f"PeerArena: a peer barrier on rank {self.rank} timed out after {self.TIMEOUT_MS} ms — a peer is stalled or "
                               "gone; results gathered since then are incomplete")

    def barrier(self):
        """Stream-ordered flag barrier over peer memory (one tiny kernel; no NCCL).  The barrier count lives on the device (epoch 0 of
        bq_peer_barrier_ex: word BQ_PEER_FLAG_EPOCH of the own flag block, advanced by the kernel), so the launch carries no per-call
        host state and a forward that contains barriers can be captured in a CUDA graph and replayed — every rank must then replay
        the same number of barriers."""
        self._raise_if_timed_out()
        self.epoch += 1                      # informational (number of barriers issued from the host)
        if self.world == 1:
            return
        L = self._L
        L.check(self.lib.bq_peer_barrier_ex(self._signals, self.rank, self.world, 0, self.TIMEOUT_MS,
                                            self._host_flag.data_ptr(), L.stream_ptr(self.device)), "bq_peer_barrier_ex")

    def push(self, local: torch.Tensor, bases, col_offset_bytes: int):
        """Copy the strided 2-D slab `local` (a column range of the tensor take() returned) to the same place in every peer's slot."""
        if self.world == 1:
            return
        L = self._L
        rows, cols = local.shape
        row_bytes = cols * local.element_size()
        dst = [b + col_offset_bytes for r, b in enumerate(bases) if r != self.rank]
        arr = (ctypes.c_void_p * len(dst))(*dst)
        stride = local.stride(0) * local.element_size()
        L.check(self.lib.bq_peer_push(local.data_ptr(), arr, len(dst), rows, row_bytes, stride, stride, L.stream_ptr(self.device)),
                "bq_peer_push")

    def timed_out(self) -> bool:
        """True if any barrier on this rank gave up waiting (synchronises the device)."""
        return bool(self.buf[:self._flag_bytes].view(torch.int32)[32].item()) or int(self._host_flag[0]) != 0

    def close(self):
        for key in self._bases:
            _ipc_release(self.lib, key)
        self._bases = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ColumnParallelLinear(nn.Module):
    """
    Column-parallel wrapper around any quantized Linear class of QUANTIZED_MODULE_MAP.

        full = get_quantized_cls("linear", cfg)(K, N, bias=True, config=cfg)        # reference-style module
        cp = ColumnParallelLinear.from_linear(full)                                   # keeps only this rank's rows
        y = cp(x)                                                                     # == full(x), bit for bit

    `gather_output=False` returns the local slab (for a following row-independent op).
    """

    def __init__(self, local: nn.Linear, out_features: int, group=None, gather_output: bool = True,
                 arena: Optional[PeerArena] = None):
        super().__init__()
        self.local = local
        self.in_features = local.in_features
        self.out_features = out_features
        self.group = group
        self.gather_output = gather_output
        self.arena = arena            # set -> fused all-gather (epilogue stores into the peers), else NCCL all-gather

    @classmethod
    def from_linear(cls, linear: nn.Linear, group=None, gather_output: bool = True, world: Optional[int] = None,
                    rank: Optional[int] = None, arena: Optional[PeerArena] = None):
        w_, r_ = _world(group)
        world = w_ if world is None else world
        rank = r_ if rank is None else rank
        lo, hi = shard_range(linear.out_features, world, rank)
        config = getattr(linear, "config", None)
        kind = type(linear)
        has_bias = linear.bias is not None
        if config is not None:
            local = kind(linear.in_features, hi - lo, bias=has_bias, config=config)
        else:
            local = kind(linear.in_features, hi - lo, bias=has_bias)
        local = local.to(linear.weight.device)
        # block_log replaces the maximum of all-zero blocks by the TENSOR-global minimum non-zero block maximum (reference
        # block_log.py:50-53): a per-shard minimum would differ, so such a module is quantised on the full tensor first and the
        # quantised values are sharded (block_fp / block_minifloat need no such care: no block crosses the cut)
        if (getattr(linear, "weight_requires_quantisation", False) and config is not None
                and config.get("name") in ("block_log", "log") and hasattr(linear, "_ensure_ptq") and linear.weight.is_cuda):
            linear._ensure_ptq()
        with torch.no_grad():
            local.weight.copy_(linear.weight[lo:hi])
            if has_bias:
                local.bias.copy_(linear.bias[lo:hi])
        # a module whose PTQ overwrite already happened holds quantised values; block_fp is NOT idempotent (a block max that
        # rounded down onto a power of two would get a smaller exponent), so carry the flag over: no second pass runs
        if hasattr(linear, "weight_requires_quantisation"):
            local.weight_requires_quantisation = linear.weight_requires_quantisation
        local.train(linear.training)
        return cls(local, linear.out_features, group=group, gather_output=gather_output, arena=arena)

    def _fused_ok(self, x) -> bool:
        loc = self.local
        return (self.arena is not None and self.gather_output and x.is_cuda and x.dtype == torch.float32
                and getattr(loc, "accepts_prequantized", lambda: False)() and loc.out_features % 32 == 0
                and getattr(loc, "_fusable", lambda _x: False)(x))

    @torch.no_grad()
    def _forward_fused(self, x: torch.Tensor) -> torch.Tensor:
        """quantize x -> GEMM whose epilogue writes this rank's slab into every rank's [M, N] buffer -> flag barrier.
        The returned tensor lives in the arena: it stays valid until `slots - 1` further calls on the same arena."""
        from .models.quantize.quantized_modules.linear import operand_format, quantize_operand_bf16

        loc, arena = self.local, self.arena
        kind, kw, block_size = operand_format(loc.config, "data_in")
        xq = quantize_operand_bf16(x, kind, kw, block_size, True).reshape(-1, self.in_features)
        m, nl = xq.shape[0], loc.out_features
        full, bases = arena.take((m, self.out_features))
        col0 = arena.rank * nl
        peers = [b + col0 * 4 for r, b in enumerate(bases) if r != arena.rank]
        loc.forward_prequantized(xq, out=full[:, col0:col0 + nl], peer_out_ptrs=peers)
        arena.barrier()
        return full.view(*x.shape[:-1], self.out_features)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self._fused_ok(x):
            return self._forward_fused(x)
        y = self.local(x)
        return all_gather_columns(y, self.group) if self.gather_output else y

    def extra_repr(self) -> str:
        world, rank = _world(self.group)
        how = "peer-store epilogue" if self.arena is not None else "nccl all-gather"
        return f"in_features={self.in_features}, out_features={self.out_features}, shard={rank}/{world}, gather={how}"


class TensorParallelOPTLayer(nn.Module):
    """
    One quantized OPT decoder layer with every Linear column-parallel over the ranks of `group` and the exchange fused into the
    producing kernels (BASELINE configs[4]: OPT-6.7B per-layer mixed-precision block_fp, column-parallel at 2 / 4 / 8 B200).

    Column-parallel only — the K-reduction of every output element stays on one rank in the single-GPU order, so the layer output
    is BIT-IDENTICAL to `OPTQuantizedDecoderLayer._fused_forward` on one GPU (row-parallel would regroup fp32 partial sums).  What
    crosses NVLink is decided by the CONSUMER of each tensor, not by the Linear that produced it:

        LN1 + x-quantizers            replicated (every rank holds h)
        q / k / v_proj                N = H/g columns = h/g whole heads per rank; epilogue applies bmm_0 / bmm_1's operand
                                      quantizers -> bf16, LOCAL only: attention is head-parallel, nothing is exchanged
        attention                     this rank's heads; epilogue applies out_proj's x-quantizer -> bf16 slab of [M, H],
                                      pushed to every peer (bq_peer_push)                                 2 B/elem on the wire
        out_proj (+ residual)         fp32 slab of h2 [M, H], stored to all ranks by the GEMM epilogue    4 B/elem (residual stream)
        LN2 + x-quantizer             replicated
        fc1 (+ ReLU + fc2's x-quantizer)  bf16 slab of [M, F], stored to all ranks by the GEMM epilogue   2 B/elem instead of 4
        fc2 (+ residual)              fp32 slab of h3 [M, H], stored to all ranks by the GEMM epilogue

    i.e. 2H + 4H + 2F + 4H = 72 KB per token for OPT-6.7B against 147 KB when each of the six Linears gathers an fp32 result.  One
    flag barrier (bq_peer_barrier) after each of the four exchanges.  `mode="nccl"` runs the same schedule with
    `all_gather_into_tensor` + permute in place of the peer stores (the library baseline).
    """

    def __init__(self, layer: nn.Module, arena: Optional[PeerArena] = None, group=None):
        super().__init__()
        world, rank = _world(group)
        at = layer.self_attn
        H, F_, heads, d = layer.embed_dim, layer.fc1.out_features, at.num_heads, at.head_dim
        if heads % world or (H // world) % 32 or (F_ // world) % 32:
            raise ValueError(f"cannot split {heads} heads / H={H} / F={F_} over {world} ranks in whole heads and multiples of 32 columns")
        self.group, self.world, self.rank, self.arena = group, world, rank, arena
        self.H, self.F, self.heads_local, self.head_dim = H, F_, heads // world, d
        self.scaling = at.scaling
        self.qc = at.quant_config
        self.plan_of = layer._fused_plan                      # format resolution stays the full layer's (same TOML node)
        self.ln1, self.ln2 = layer.self_attn_layer_norm, layer.final_layer_norm
        cut = lambda lin: ColumnParallelLinear.from_linear(lin, group=group, gather_output=False, world=world, rank=rank).local
        self.q_proj, self.k_proj, self.v_proj, self.out_proj = cut(at.q_proj), cut(at.k_proj), cut(at.v_proj), cut(at.out_proj)
        self.fc1, self.fc2 = cut(layer.fc1), cut(layer.fc2)
        self.fc1_exchange = "push"             # or "epilogue": remote stores from the fc1 GEMM epilogue (A/B)

    def nvlink_bytes_per_rank(self, tokens: int, mode: str = "fused") -> int:
        """Bytes one rank SENDS per layer call."""
        per_tok = (2 * self.H + 4 * self.H + 2 * self.F + 4 * self.H) // self.world
        return tokens * per_tok * (self.world - 1)

    @torch.no_grad()
    def forward(self, h: torch.Tensor, mode: str = "fused", events: Optional[list] = None) -> torch.Tensor:
        from .models.quantize.quantized_functions.attention import fused_causal_attention_q
        from .models.quantize.quantized_functions.fused_glue import norm_quantize

        B, S, H = h.shape
        M, g, r = B * S, self.world, self.rank
        Hl, Fl = H // g, self.F // g
        c0, f0 = r * Hl, r * Fl
        plan = self.plan_of(S)
        if plan is None or plan.get("mode") == "split":
            raise NotImplementedError("TensorParallelOPTLayer needs a layer the fused path serves (PTQ block_fp / block_minifloat, "
                                      "[1,16] blocks, <= 8 significant bits)")
        fused = mode == "fused" and self.arena is not None and g > 1
        arena = self.arena
        ln1, ln2 = self.ln1, self.ln2
        h2d = h.reshape(M, H)

        def mark(name):
            if events is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                events.append((name, e))

        mark("start")
        xq_q, xq_k, xq_v = norm_quantize(h, ln1.weight, ln1.bias, ln1.eps, [plan["q_in"], plan["k_in"], plan["v_in"]])
        mark("ln1")
        Qq = self.q_proj.forward_prequantized(xq_q, scale=self.scaling, out_format=plan["q_out"])
        Kq = self.k_proj.forward_prequantized(xq_k, out_format=plan["k_out"], out_blocks_along_rows=True)
        Vq = self.v_proj.forward_prequantized(xq_v, out_format=plan["v_out"])
        mark("qkv")
        # attention on this rank's heads; the bf16 result (already in out_proj's x-format) lands in the gathered [M, H] buffer
        if fused:
            attn_full, bases = arena.take((M, H), torch.bfloat16)
            slab = attn_full[:, c0:c0 + Hl]
            fused_causal_attention_q(Qq, Kq, Vq, self.qc["bmm_1"], self.heads_local, B, S, 1.0, out_cfg=self.out_proj.config, out=slab)
            mark("attention")
            arena.push(slab, bases, c0 * 2)
            arena.barrier()
        else:
            o_l = fused_causal_attention_q(Qq, Kq, Vq, self.qc["bmm_1"], self.heads_local, B, S, 1.0, out_cfg=self.out_proj.config)
            mark("attention")
            attn_full = all_gather_columns(o_l.view(M, Hl), self.group)
        mark("gather_attn")
        if fused:
            h2_full, bases = arena.take((M, H), torch.float32)
            self.out_proj.forward_prequantized(attn_full, residual=h2d[:, c0:c0 + Hl], out=h2_full[:, c0:c0 + Hl],
                                               peer_out_ptrs=[b + c0 * 4 for i, b in enumerate(bases) if i != r])
            arena.barrier()
        else:
            h2_l = self.out_proj.forward_prequantized(attn_full, residual=h2d[:, c0:c0 + Hl])
            h2_full = all_gather_columns(h2_l, self.group)
        mark("out_proj")
        (x1,) = norm_quantize(h2_full, ln2.weight, ln2.bias, ln2.eps, [plan["fc1_in"]])
        mark("ln2")
        if fused:
            a_full, bases = arena.take((M, self.F), torch.bfloat16)
            if self.fc1_exchange == "epilogue":
                self.fc1.forward_prequantized(x1, relu=True, out_format=plan["fc2_in"], out=a_full[:, f0:f0 + Fl],
                                              peer_out_ptrs=[b + f0 * 2 for i, b in enumerate(bases) if i != r])
            else:
                # bf16 slab: the epilogue's remote stores are 64-byte row segments (32 columns x 2 bytes per warp row) and sustain
                # ~400 GB/s; the push kernel's 512-byte warp stores reach 560-650 GB/s (profiles/r02_peer_exchange_n8.json) — store
                # locally, then push: fc1 0.34 -> 0.26 ms at 4096 tokens on 8 GPUs
                slab = a_full[:, f0:f0 + Fl]
                self.fc1.forward_prequantized(x1, relu=True, out_format=plan["fc2_in"], out=slab)
                arena.push(slab, bases, f0 * 2)
            arena.barrier()
        else:
            a_l = self.fc1.forward_prequantized(x1, relu=True, out_format=plan["fc2_in"])
            a_full = all_gather_columns(a_l, self.group)
        mark("fc1")
        if fused:
            h3_full, bases = arena.take((M, H), torch.float32)
            self.fc2.forward_prequantized(a_full, residual=h2_full[:, c0:c0 + Hl], out=h3_full[:, c0:c0 + Hl],
                                          peer_out_ptrs=[b + c0 * 4 for i, b in enumerate(bases) if i != r])
            arena.barrier()
        else:
            h3_l = self.fc2.forward_prequantized(a_full, residual=h2_full[:, c0:c0 + Hl])
            h3_full = all_gather_columns(h3_l, self.group)
        mark("fc2")
        return h3_full.view(B, S, H)


def column_parallelize(model: nn.Module, names: Iterable[str], group=None) -> List[str]:
    """Replace the named quantized Linear submodules of `model` (e.g. "model.decoder.layers.0.fc1") in place."""
    done = []
    for name in names:
        parent_name, _, leaf = name.rpartition(".")
        parent = model.get_submodule(parent_name) if parent_name else model
        setattr(parent, leaf, ColumnParallelLinear.from_linear(getattr(parent, leaf), group=group))
        done.append(name)
    return done
