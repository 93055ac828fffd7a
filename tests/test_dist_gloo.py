"""world_size-2 gloo tests (CPU) of the multi-GPU host logic in llm_mixed_q_b200/dist.py.

The CUDA Linear modules cannot run here (no GPU, no CPU fallback by design), so the per-shard compute is a
test-local nn.Linear subclass backed by the ORACLE (oracle/ is test infrastructure); what is under test is the
sharding, the gather layout and the perplexity reduction."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from llm_mixed_q_b200 import dist as D  # noqa: E402
from oracle import oracle as O  # noqa: E402

CFG = {"name": "block_fp", "bypass": False, "is_ptq": True}
for _p in ("data_in", "weight", "bias"):
    CFG.update({f"{_p}_width": 6, f"{_p}_exponent_width": 8, f"{_p}_exponent_bias": 127,
                f"{_p}_block_size": [16] if _p == "bias" else [1, 16]})


class OracleLinear(nn.Linear):
    """Reference-shaped quantized Linear whose forward is the oracle's PTQ path."""

    def __init__(self, in_features, out_features, bias=True, config=None):
        super().__init__(in_features, out_features, bias)
        self.config = config
        self.weight_requires_quantisation = True

    def forward(self, x):
        with torch.no_grad():
            y, wq, bq = O.linear_forward(x, self.weight.detach(), None if self.bias is None else self.bias.detach(), self.config)
            self.last_wq, self.last_bq = wq, bq
        return y


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _full_linear(K=64, N=96):
    g = torch.Generator().manual_seed(7)
    lin = OracleLinear(K, N, bias=True, config=CFG)
    with torch.no_grad():
        lin.weight.copy_(torch.randn(N, K, generator=g) * 0.05)
        lin.bias.copy_(torch.randn(N, generator=g) * 0.05)
    x = torch.randn(2, 5, K, generator=g)
    return lin, x


class _FakeLM(nn.Module):
    """loss = mean of input_ids (deterministic per batch) — enough to test the reduction."""

    def __init__(self):
        super().__init__()
        self.p = nn.Parameter(torch.zeros(1))

    def forward(self, input_ids=None, labels=None):
        class Out:
            pass
        o = Out()
        o.loss = input_ids.float().mean() / 100.0
        return o


def _batches():
    g = torch.Generator().manual_seed(3)
    return [{"input_ids": torch.randint(0, 100, (3, 8), generator=g), "labels": None} for _ in range(5)]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        lin, x = _full_linear()
        cp = D.ColumnParallelLinear.from_linear(lin)
        y = cp(x)
        lo, hi = D.shard_range(lin.out_features, world, rank)
        ppl = D.dp_perplexity(_FakeLM(), _batches())
        q.put((rank, y, cp.local.last_wq, cp.local.last_bq, lo, hi, ppl))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(420)
def test_column_parallel_and_dp_perplexity_world2():
    world = 2
    ctx = mp.get_context("spawn")
    results, last_err = None, None
    for attempt in range(3):                  # a rendezvous can lose its port to another process between _free_port() and bind
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
        for p in procs:
            p.start()
        try:
            got = [q.get(timeout=100) for _ in range(world)]
            for p in procs:
                p.join(timeout=30)
            if all(p.exitcode == 0 for p in procs):
                results = got
                break
            last_err = f"exit codes {[p.exitcode for p in procs]}"
        except Exception as e:                # queue.Empty: a worker died before reporting
            last_err = repr(e)
        for p in procs:
            if p.is_alive():
                p.kill()
            p.join(timeout=10)
    assert results is not None, last_err
    lin, x = _full_linear()
    y_full = lin(x)
    ref_ppl = D.dp_perplexity(_FakeLM(), _batches())        # single process: the reference's loop
    for rank, y, wq, bq, lo, hi, ppl in results:
        # shards quantise to exactly the rows of the fully-quantised weight / bias (no block crosses the cut)
        assert torch.equal(wq.view(torch.int32), lin.last_wq[lo:hi].view(torch.int32))
        assert torch.equal(bq.view(torch.int32), lin.last_bq[lo:hi].view(torch.int32))
        assert y.shape == y_full.shape
        # CPU sgemm may block the K-sum differently for different N; the CUDA kernel does not (tests -m gpu)
        torch.testing.assert_close(y, y_full, rtol=1e-6, atol=1e-6)
        assert ppl["num_samples"] == ref_ppl["num_samples"] == 15
        assert ppl["seq_len"] == 8
        assert abs(ppl["loss"] - ref_ppl["loss"]) < 1e-12
        assert abs(ppl["perplexity"] - ref_ppl["perplexity"]) < 1e-9


def test_shard_range_rules():
    assert D.shard_range(4096, 8, 3) == (1536, 2048)
    assert D.shard_range(16384, 8, 7) == (14336, 16384)
    with pytest.raises(ValueError):
        D.shard_range(4096 + 16, 8, 0)          # not divisible
    with pytest.raises(ValueError):
        D.shard_range(96, 4, 0)                 # 24 rows per rank: a 16-wide bias block would straddle the cut
    assert list(D.shard_batches(5, 2, 1)) == [1, 3]


def test_perplexity_matches_oracle_reduction():
    losses = [2.0, 4.0, 3.0]
    ref = O.perplexity_from_losses(losses, batch_size=3, seq_len=7)
    got = D.perplexity_from_sums(sum(l * 3 * 7 for l in losses), 9, 7)
    assert abs(got["perplexity"] - ref) < 1e-9 * ref
