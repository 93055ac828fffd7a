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
   values to quantising the full matrix, because no block crosses the cut), computes its column slab
   of y from the replicated x, and the slabs are all-gathered.  The K-reduction of every output element is
   done by the same kernel in the same order as on one GPU, so the result is bit-identical to 1 GPU.
   `ColumnParallelLinear`.
"""
from __future__ import annotations

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


class ColumnParallelLinear(nn.Module):
    """
    Column-parallel wrapper around any quantized Linear class of QUANTIZED_MODULE_MAP.

        full = get_quantized_cls("linear", cfg)(K, N, bias=True, config=cfg)        # reference-style module
        cp = ColumnParallelLinear.from_linear(full)                                   # keeps only this rank's rows
        y = cp(x)                                                                     # == full(x), bit for bit

    `gather_output=False` returns the local slab (for a following row-independent op).
    """

    def __init__(self, local: nn.Linear, out_features: int, group=None, gather_output: bool = True):
        super().__init__()
        self.local = local
        self.in_features = local.in_features
        self.out_features = out_features
        self.group = group
        self.gather_output = gather_output

    @classmethod
    def from_linear(cls, linear: nn.Linear, group=None, gather_output: bool = True, world: Optional[int] = None,
                    rank: Optional[int] = None):
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
        with torch.no_grad():
            local.weight.copy_(linear.weight[lo:hi])
            if has_bias:
                local.bias.copy_(linear.bias[lo:hi])
        # a module whose PTQ overwrite already happened holds quantised values; block_fp is NOT idempotent (a block max that
        # rounded down onto a power of two would get a smaller exponent), so carry the flag over: no second pass runs
        if hasattr(linear, "weight_requires_quantisation"):
            local.weight_requires_quantisation = linear.weight_requires_quantisation
        local.train(linear.training)
        return cls(local, linear.out_features, group=group, gather_output=gather_output)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        y = self.local(x)
        return all_gather_columns(y, self.group) if self.gather_output else y

    def extra_repr(self) -> str:
        world, rank = _world(self.group)
        return f"in_features={self.in_features}, out_features={self.out_features}, shard={rank}/{world}"


def column_parallelize(model: nn.Module, names: Iterable[str], group=None) -> List[str]:
    """Replace the named quantized Linear submodules of `model` (e.g. "model.decoder.layers.0.fc1") in place."""
    done = []
    for name in names:
        parent_name, _, leaf = name.rpartition(".")
        parent = model.get_submodule(parent_name) if parent_name else model
        setattr(parent, leaf, ColumnParallelLinear.from_linear(getattr(parent, leaf), group=group))
        done.append(name)
    return done
