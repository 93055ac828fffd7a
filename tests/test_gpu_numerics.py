"""GPU: the building blocks the bit-exactness argument rests on (SURVEY.md §7 on-device checklist)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_log2_shortcuts_match_libdevice_on_every_positive_float():
    from llm_mixed_q_b200 import _lib as L

    lib = L.load()
    out = torch.full((3,), -1, dtype=torch.int64, device="cuda")
    L.check(lib.bq_selftest_log2(out.data_ptr(), L.stream_ptr()), "bq_selftest_log2")
    torch.cuda.synchronize()
    assert out.tolist() == [0, 0, 0], f"ceil/floor/rint(log2f) shortcut mismatches: {out.tolist()}"


def test_pow2_is_exact_on_torch_cuda():
    e = torch.arange(-160, 131, dtype=torch.float32, device="cuda")
    got = (2 ** e).cpu()
    exp = torch.tensor([2.0 ** int(k) if -149 <= k <= 127 else (float("inf") if k > 127 else 0.0) for k in range(-160, 131)],
                       dtype=torch.float64).to(torch.float32)
    assert torch.equal(got.view(torch.int32), exp.view(torch.int32))


def test_torch_cuda_log2_matches_libdevice_where_it_matters():
    """torch.log2 on CUDA and the kernels' log2f must agree after ceil/floor/round on the cliff neighbourhoods:
    checked indirectly by the bit-exact quantizer tests; here: CPU vs CUDA oracle agreement on the cliffs."""
    k = torch.arange(-60, 61, dtype=torch.float32)
    d = torch.arange(-16, 33, dtype=torch.int32)
    for base in (1.0, 2.0 ** 0.5):
        b = torch.tensor(base, dtype=torch.float32).view(torch.int32)
        vals = ((b + d).view(torch.float32)[None, :] * torch.exp2(k)[:, None]).reshape(-1)
        lc, lg = torch.log2(vals), torch.log2(vals.cuda()).cpu()
        n = vals.numel()
        # informational only: CPU (SLEEF) and CUDA (libdevice) may differ here; record how often
        diffs = {name: int((f(lc) != f(lg)).sum()) for name, f in (("ceil", torch.ceil), ("floor", torch.floor), ("round", torch.round))}
        print(f"log2 cliff CPU-vs-CUDA disagreements at base {base:.3f}: {diffs} of {n}")
