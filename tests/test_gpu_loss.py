"""GPU parity of the shifted token cross-entropy: bq_token_ce_mean against the reference's own loss tail
(modeling_opt.py:1086-1098: shifted copy + CrossEntropyLoss) evaluated by torch in fp64 and fp32 on the same logits.

Stated tolerance: the kernel's log-sum-exp uses ex2.approx on fma-scaled arguments and a different summation order;
|loss - fp64 loss| <= 2e-6 * max(1, |loss|) — the same order as torch's own fp32 CrossEntropyLoss error, which is printed
beside it.  ignore_index rows are excluded from sum and count exactly like reduction="mean"."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def ref_loss(logits, labels, shift, dtype):
    lg = logits.to(dtype)
    if shift:
        lg, lb = lg[..., :-1, :].contiguous(), labels[..., 1:].contiguous()
    else:
        lb = labels
    return F.cross_entropy(lg.view(-1, lg.shape[-1]), lb.reshape(-1))


@pytest.mark.parametrize("B,S,V,scale", [(2, 64, 50272, 1.0), (3, 17, 32000, 8.0), (1, 5, 1003, 3.0), (4, 33, 30, 20.0), (2, 2, 4, 1.0)])
@pytest.mark.parametrize("shift", [True, False])
def test_token_ce_matches_torch(B, S, V, scale, shift):
    from llm_mixed_q_b200.models.quantize.quantized_functions.loss import causal_lm_loss

    g = torch.Generator(device="cuda").manual_seed(B * 1000 + S + V)
    logits = torch.randn(B, S, V, device="cuda", generator=g) * scale
    labels = torch.randint(0, V, (B, S), device="cuda", generator=g)
    got = float(causal_lm_loss(logits, labels, shift=shift))
    want64 = float(ref_loss(logits, labels, shift, torch.float64))
    want32 = float(ref_loss(logits, labels, shift, torch.float32))
    tol = 2e-6 * max(1.0, abs(want64))
    assert abs(got - want64) <= tol, (got, want64, want32)
    assert abs(got - want32) <= 2 * tol


def test_token_ce_ignore_index_and_strides():
    from llm_mixed_q_b200.models.quantize.quantized_functions.loss import causal_lm_loss

    g = torch.Generator(device="cuda").manual_seed(7)
    B, S, V = 3, 40, 2050                                   # V % 4 != 0 -> rows not 16-byte aligned
    big = torch.randn(B, S, V + 6, device="cuda", generator=g) * 4
    logits = big[..., :V]                                   # row stride V + 6
    labels = torch.randint(0, V, (B, S), device="cuda", generator=g)
    labels[0, 3:9] = -100
    labels[2, :] = -100
    got = float(causal_lm_loss(logits, labels, shift=True))
    want = float(ref_loss(logits, labels, True, torch.float64))
    assert abs(got - want) <= 2e-6 * max(1.0, abs(want))
    # no valid row -> NaN like torch
    labels[:] = -100
    assert torch.isnan(causal_lm_loss(logits, labels, shift=True))
    # deterministic
    labels = torch.randint(0, V, (B, S), device="cuda", generator=g)
    a = causal_lm_loss(logits, labels)
    b = causal_lm_loss(logits, labels)
    assert torch.equal(a, b)


def test_token_ce_full_size_property():
    """BASELINE size (8 x 2048 x 50272): adding a per-row constant to the logits leaves the loss unchanged (shift invariance
    of log-softmax), and the loss equals torch's within the stated tolerance."""
    from llm_mixed_q_b200.models.quantize.quantized_functions.loss import causal_lm_loss

    g = torch.Generator(device="cuda").manual_seed(11)
    B, S, V = 8, 2048, 50272
    logits = torch.randn(B, S, V, device="cuda", generator=g)
    labels = torch.randint(0, V, (B, S), device="cuda", generator=g)
    got = float(causal_lm_loss(logits, labels))
    want = float(F.cross_entropy(logits[:, :-1].reshape(-1, V), labels[:, 1:].reshape(-1)))
    assert abs(got - want) <= 4e-6 * abs(want)
    logits += torch.randn(B, S, 1, device="cuda", generator=g) * 4
    got2 = float(causal_lm_loss(logits, labels))
    assert abs(got2 - got) <= 4e-6 * abs(got)


def test_token_ce_rejects_cpu_tensors():
    from llm_mixed_q_b200.models.quantize.quantized_functions.loss import causal_lm_loss

    with pytest.raises(RuntimeError):
        causal_lm_loss(torch.randn(1, 4, 8), torch.zeros(1, 4, dtype=torch.long))
