"""GPU end-to-end parity: tiny random-init OPT / Llama with the reference's weights (golden state dicts) through
our quantized module classes vs the reference's own CPU forward (logits + loss).

Rounding makes quantisation discontinuous: an ulp-level difference upstream (fp32 accumulation order in a GEMM)
can flip an element by one quantisation step downstream (SURVEY.md hard part 6), so logits are compared
statistically: loss within 2e-3 relative, logits max-abs error small against their spread."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLD

pytestmark = pytest.mark.gpu


def load(tag):
    return np.load(os.path.join(GOLD, tag + ".npz"))


def raw_configs():
    with open(os.path.join(GOLD, "configs.json")) as f:
        g = json.load(f)
    return g


def clone(d):
    return json.loads(json.dumps(d))


@pytest.mark.parametrize("tag,cfgkey", [("opt_tiny_bfp6", "bfp_6bit.toml"), ("opt_tiny_bfp4", "bfp_4bit.toml"),
                                        ("opt_tiny_mixed", None)])
def test_opt_tiny_matches_reference_forward(tag, cfgkey):
    from llm_mixed_q_b200.models.opt_quantized import OPTQuantizedConfig, OPTQuantizedForCausalLM

    g = raw_configs()
    qc = clone(g["raw"][cfgkey]) if cfgkey else clone(g["mixed_raw"])
    z = load(tag)
    cfg = OPTQuantizedConfig(hidden_size=64, num_hidden_layers=2, ffn_dim=128, num_attention_heads=4, vocab_size=512,
                             max_position_embeddings=64, quant_config=qc,
                             tie_word_embeddings=False)   # the golden run (transformers 5.5) had an untied, separately initialised lm_head
    model = OPTQuantizedForCausalLM(cfg).eval()
    sd = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd::")}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not missing, missing
    model = model.cuda()
    ids = torch.from_numpy(z["input_ids"]).cuda()
    with torch.no_grad():
        out = model(input_ids=ids, labels=ids)
    ref_logits = torch.from_numpy(z["logits"])
    ref_loss = float(z["loss"])
    assert abs(float(out.loss) - ref_loss) <= 2e-3 * abs(ref_loss), (float(out.loss), ref_loss)
    err = (out.logits.cpu() - ref_logits).abs()
    spread = float(ref_logits.std())
    assert float(err.mean()) <= 0.02 * spread and float(err.max()) <= 0.5 * spread, (float(err.mean()), float(err.max()), spread)


@pytest.mark.parametrize("tag,cfgkey", [("llama_tiny_bfp6", "bfp_6bit.toml"), ("llama_tiny_bmf8", "block_minifloat.toml"),
                                        ("llama_tiny_bl8", "block_log.toml")])
def test_llama_tiny_matches_reference_forward(tag, cfgkey):
    from llm_mixed_q_b200.models.llama_quantized import LlamaQuantizedConfig, LlamaQuantizedForCausalLM

    g = raw_configs()
    z = load(tag)
    cfg = LlamaQuantizedConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                               vocab_size=512, max_position_embeddings=64, initializer_range=float(z["init"]),
                               quant_config=clone(g["raw"][cfgkey]))
    model = LlamaQuantizedForCausalLM(cfg).eval()
    sd = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd::")}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not missing, missing
    model = model.cuda()
    ids = torch.from_numpy(z["input_ids"]).cuda()
    with torch.no_grad():
        out = model(input_ids=ids, labels=ids)
    ref_logits = torch.from_numpy(z["logits"])
    ref_loss = float(z["loss"])
    assert abs(float(out.loss) - ref_loss) <= 5e-3 * abs(ref_loss), (float(out.loss), ref_loss)
    err = (out.logits.cpu() - ref_logits).abs()
    spread = float(ref_logits.std())
    assert float(err.mean()) <= 0.03 * spread, (float(err.mean()), float(err.max()), spread)


@pytest.mark.parametrize("tag,cfgkey", [("llama_small_bfp6", "bfp_6bit.toml"), ("llama_small_bmf8", "block_minifloat.toml")])
def test_llama_fused_layers_match_reference_forward(tag, cfgkey):
    """The FUSED Llama layer at head_dim 128 — RMSNorm + quantize, q | k | v as one GEMM with the RoPE / operand-quantizer epilogue,
    one-kernel attention (key-padding bitmap for the right-padded row), o_proj / down_proj residual epilogues, gate | up as one GEMM
    with the gated-SiLU epilogue — against the unmodified reference's forward (oracle/gen_golden_llama_fused.py; reference
    models/llama_quantized/modeling_llama.py:246, :274-344).  The op-by-op path of this package is held to the same golden, and the
    fused path may not be further from the reference than it (beyond the statistical noise of one-step rounding flips)."""
    from llm_mixed_q_b200 import _lib as L
    from llm_mixed_q_b200.models.llama_quantized import LlamaQuantizedConfig, LlamaQuantizedForCausalLM

    g = raw_configs()
    z = load(tag)
    sd = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd::")}
    ids, am, labels = (torch.from_numpy(z[k]).cuda() for k in ("input_ids", "attention_mask", "labels"))
    valid = am.bool().cpu()
    res = {}
    for fused in (True, False):
        cfg = LlamaQuantizedConfig(hidden_size=256, intermediate_size=352, num_hidden_layers=2, num_attention_heads=2, vocab_size=512,
                                   max_position_embeddings=128, initializer_range=float(z["initializer_range"]), pad_token_id=0,
                                   quant_config=clone(g["raw"][cfgkey]))
        model = LlamaQuantizedForCausalLM(cfg).eval()
        missing, _ = model.load_state_dict(sd, strict=False)
        assert not missing, missing
        model = model.cuda()
        model.model.fused_glue = fused
        n_attn, n_epi = L.launch_counts()["attention_causal_kernel"], L.launch_counts()["gemm_bf16_tn_kernel<epilogue>"]
        with torch.no_grad():
            out = model(input_ids=ids, attention_mask=am, labels=labels)
            out1 = model(input_ids=ids[:1], labels=ids[:1])
        if fused:
            layer = model.model.layers[0]
            assert layer._fused_plan(ids.shape[1]) is not None and layer._fused_plan(ids.shape[1]).get("mode") != "split"
            assert L.launch_counts()["attention_causal_kernel"] - n_attn == 4                 # 2 layers x 2 forwards, one kernel each
            # per layer and forward: q | k | v (1), o_proj (1), gate | up (1), down_proj (1)
            assert L.launch_counts()["gemm_bf16_tn_kernel<epilogue>"] - n_epi == 16
            assert getattr(layer.self_attn.q_proj, "_qkv_cache", None) is not None and getattr(layer.mlp.gate_proj, "_gu_cache", None) is not None
        for o, key, sel in ((out, "", valid), (out1, "_unpadded_row0", torch.ones(1, ids.shape[1], dtype=torch.bool))):
            ref_logits, ref_loss = torch.from_numpy(z["logits" + key]), float(z["loss" + key])
            assert abs(float(o.loss) - ref_loss) <= 5e-3 * abs(ref_loss), (fused, key, float(o.loss), ref_loss)
            err = (o.logits.cpu() - ref_logits).abs()[sel]
            spread = float(ref_logits[sel].std())
            res[(fused, key)] = float(err.mean()) / spread
            assert float(err.mean()) <= 0.03 * spread, (fused, key, float(err.mean()), float(err.max()), spread)
    assert res[(True, "")] <= 1.5 * res[(False, "")] + 2e-3, res


def _opt125m_from_seed():
    """BASELINE configs[0] model: OPT-125M shape (the config class defaults), random init under seed 0 on the CPU — bit-identical
    to the reference model the golden was generated from (oracle/gen_golden_opt125m.py), which the checksums re-verify."""
    from llm_mixed_q_b200.models.opt_quantized import OPTQuantizedConfig, OPTQuantizedForCausalLM

    z = load("opt125m_bfp6")
    torch.manual_seed(0)
    model = OPTQuantizedForCausalLM(OPTQuantizedConfig(quant_config=clone(raw_configs()["raw"]["bfp_6bit.toml"]),
                                                       tie_word_embeddings=False)).eval()
    sd = model.state_dict()
    for k, (s, a) in zip(z["checksum_keys"], z["checksum_vals"]):
        v = sd[str(k)].double()
        # float64 sums: exact up to summation order (threads / vector width differ between hosts)
        assert abs(float(v.sum()) - s) <= 1e-9 * (a + 1e-30) and abs(float(v.abs().sum()) - a) <= 1e-9 * a, \
            f"seeded init of {k} differs from the reference's"
    return model, z


@pytest.mark.timeout(600)
def test_opt125m_config1_fused_and_op_by_op_match_reference_forward():
    """BASELINE configs[0] at FULL size (1 x 2048 tokens, OPT-125M, W6A6 block_fp on every Linear and both bmms): loss, per-token
    log-partition and a 128 x 786 sample of the logits of the unmodified reference's CPU forward against (a) the fused layers,
    (b) the same modules op by op.  Perplexity = exp(loss) agrees to 2e-3 relative in the loss."""
    model, z = _opt125m_from_seed()
    dec = model.model.decoder
    sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}        # PTQ overwrites the parameters on the first forward
    ids = torch.from_numpy(z["input_ids"]).cuda()
    ref_loss, ref_sub, ref_lse, spread = float(z["loss"]), torch.from_numpy(z["logits_sub"]), torch.from_numpy(z["logits_row_lse"]), float(z["logits_std"])
    model = model.cuda()
    results = {}
    for name, fused in (("fused", True), ("op-by-op", False)):
        model.load_state_dict(sd0)
        for lyr in dec.layers:
            for m in lyr.modules():
                if hasattr(m, "weight_requires_quantisation"):
                    m.weight_requires_quantisation = True
        dec.fused_glue, dec.fused_attention = fused, fused
        if fused:
            assert dec.layers[0]._fused_plan(ids.shape[1]) is not None
        with torch.no_grad():
            out = model(input_ids=ids, labels=ids)
        logits = out.logits[0]
        err = (logits[::16, ::64].cpu() - ref_sub).abs()
        lse = torch.logsumexp(logits.double(), -1).cpu()
        results[name] = (float(out.loss), float(err.mean()), float(err.max()), float((lse - ref_lse).abs().max()))
    dec.fused_glue, dec.fused_attention = True, True
    # noise-floor control: the oracle port on this GPU (torch-CUDA fp32, cuBLAS) against the same CPU golden — how far a CORRECT
    # implementation with another GEMM accumulation order drifts.  The CUDA path is held to 1.25x of it (see test_gpu_parity_opt13b)
    from llm_mixed_q_b200.models.opt_quantized import parse_opt_quantized_config
    from oracle import opt_ref

    torch.backends.cuda.matmul.allow_tf32 = False
    qc = parse_opt_quantized_config(clone(raw_configs()["raw"]["bfp_6bit.toml"]), 12)
    with torch.no_grad():
        lg_c, loss_c = opt_ref.opt_forward({k: v.cuda() for k, v in sd0.items()}, qc, ids, 12, 12, labels=ids)
    err_c = (lg_c[0][::16, ::64].cpu() - ref_sub).abs()
    floor_mean, floor_max = float(err_c.mean()), float(err_c.max())
    print("opt125m config-1 parity (loss, mean|dlogit|, max|dlogit|, max|dLSE|):", results, "reference loss", ref_loss, "logit std", spread,
          "control (oracle on GPU vs CPU golden): loss", float(loss_c), "mean|dlogit|", floor_mean, "max", floor_max)
    for name, (loss, e_mean, e_max, d_lse) in results.items():
        assert e_mean <= 1.25 * floor_mean + 1e-3 * spread, (name, e_mean, floor_mean)
        # 12 layers of 6-bit rounding: an ulp-level difference in a GEMM's accumulation order flips individual elements by one
        # quantisation step and the flips diffuse through the residual stream — individual logits move by a few % of their
        # spread (in BOTH paths, against a reference that itself differs from run to run on another BLAS), while the loss
        # (= log perplexity) and the per-token log-partition agree to 1e-4
        assert abs(loss - ref_loss) <= 2e-4 * abs(ref_loss), (name, loss, ref_loss)
        assert e_mean <= 0.08 * spread and e_max <= 0.5 * spread, (name, e_mean, e_max, spread)
        assert d_lse <= 0.01, (name, d_lse)
    # the fused layers must not be further from the reference than the op-by-op path (bit-exact quantizers + same GEMM kernel) is
    assert results["fused"][1] <= 1.5 * results["op-by-op"][1] + 1e-3, results


@pytest.mark.timeout(600)
def test_opt_1p3b_headline_shape_fused_matches_op_by_op():
    """BASELINE configs[2] shape (OPT-1.3B, W6A6 block_fp on every Linear and both bmms, seq 2048; batch 2 to bound the op-by-op
    path's 4 GB of fp32 scores per layer): the 8-kernel fused layers against the same modules run op by op (bit-exact quantizer
    kernels, one GEMM kernel, torch softmax) on the synthetic token stream bench.py uses — loss to 2e-4 relative."""
    import bench

    model = bench.build_model(torch.device("cuda", 0))
    dec = model.model.decoder
    ids = torch.randint(0, bench.OPT13B["vocab_size"], (2, bench.SEQ), generator=torch.Generator().manual_seed(0)).cuda()
    with torch.no_grad():
        assert dec.layers[0]._fused_plan(bench.SEQ) is not None
        fused = model(input_ids=ids, labels=ids)                       # also performs the one-off PTQ weight overwrite
        lf, logits_f = float(fused.loss), fused.logits[:, ::64, ::97].float().cpu()
        del fused
        dec.fused_glue, dec.fused_attention = False, False
        ref = model(input_ids=ids, labels=ids)
        lr, logits_r = float(ref.loss), ref.logits[:, ::64, ::97].float().cpu()
        spread = float(ref.logits.std())
        del ref
    dec.fused_glue, dec.fused_attention = True, True
    err = (logits_f - logits_r).abs()
    print("opt-1.3b fused vs op-by-op: loss", lf, lr, "mean|dlogit|", float(err.mean()), "max", float(err.max()), "logit std", spread)
    assert abs(lf - lr) <= 2e-4 * abs(lr), (lf, lr)
    assert float(err.mean()) <= 0.08 * spread and float(err.max()) <= 0.6 * spread, (float(err.mean()), float(err.max()), spread)


def _attention_launches():
    from llm_mixed_q_b200 import _lib as L

    return L.launch_counts()["attention_causal_kernel"]


def test_opt_padded_batch_runs_on_the_fused_kernels_and_matches_reference_forward():
    """A right-padded batch (reference mask path opt_quantized/modeling_opt.py:520-548) stays on the fused layers: the attention
    kernel takes the key-padding bitmap.  Golden = the unmodified reference's forward (oracle/gen_golden_masked_attention.py);
    the op-by-op path of this package is the second witness."""
    from llm_mixed_q_b200.models.opt_quantized import OPTQuantizedConfig, OPTQuantizedForCausalLM

    z = load("opt_small_padded_bfp6")
    qc = clone(raw_configs()["raw"]["bfp_6bit.toml"])
    cfg = OPTQuantizedConfig(hidden_size=128, num_hidden_layers=2, ffn_dim=256, num_attention_heads=2, vocab_size=512,
                             max_position_embeddings=128, quant_config=qc, pad_token_id=1, tie_word_embeddings=False)
    sd = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd::")}
    ids, am, labels = (torch.from_numpy(z[k]).cuda() for k in ("input_ids", "attention_mask", "labels"))
    ref_logits, ref_loss = torch.from_numpy(z["logits"]), float(z["loss"])
    valid = am.bool().cpu()
    outs = {}
    for fused in (True, False):
        model = OPTQuantizedForCausalLM(cfg).eval()
        missing, _ = model.load_state_dict(sd, strict=False)
        assert not missing, missing
        model = model.cuda()
        model.model.decoder.fused_attention = fused
        model.model.decoder.fused_glue = fused
        n0 = _attention_launches()
        with torch.no_grad():
            out = model(input_ids=ids, attention_mask=am, labels=labels)
        assert _attention_launches() - n0 == (2 if fused else 0)
        outs[fused] = out
        assert abs(float(out.loss) - ref_loss) <= 2e-3 * abs(ref_loss), (fused, float(out.loss), ref_loss)
        # logits of real tokens (pad-token rows are computed too but carry no label)
        err = (out.logits.cpu() - ref_logits).abs()[valid]
        spread = float(ref_logits[valid].std())
        assert float(err.mean()) <= 0.02 * spread and float(err.max()) <= 0.5 * spread, (fused, float(err.mean()), float(err.max()), spread)
    # left padding: fully masked causal rows -> the kernel is not used (the reference's uniform-over-all-keys rows)
    am_left = torch.flip(am, dims=[1])
    model.model.decoder.fused_attention = True
    model.model.decoder.fused_glue = True
    n0 = _attention_launches()
    with torch.no_grad():
        model(input_ids=ids, attention_mask=am_left)
    assert _attention_launches() == n0


def test_bert_head_dim_64_runs_bidirectional_fused_attention_and_matches_reference_forward():
    """BERT (bidirectional attention + key-padding mask, bert_quantized/modeling_bert.py:366-435) through bq_attention_masked
    against the unmodified reference's forward, with the op-by-op path as second witness."""
    from llm_mixed_q_b200.models.bert_quantized import BertQuantizedConfig, BertQuantizedForSequenceClassification

    z = load("bert_small_bfp6")
    qc = clone(raw_configs()["raw"]["bfp_6bit.toml"])
    cfg = BertQuantizedConfig(quant_config=qc, initializer_range=0.05, vocab_size=512, hidden_size=128, num_hidden_layers=2,
                              num_attention_heads=2, intermediate_size=256, max_position_embeddings=128, num_labels=3)
    sd = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd::")}
    ids, am, tt = (torch.from_numpy(z[k]).cuda() for k in ("input_ids", "attention_mask", "token_type_ids"))
    ref_h, ref_logits = torch.from_numpy(z["last_hidden"]), torch.from_numpy(z["logits"])
    valid = am.bool().cpu()
    res = {}
    for fused in (True, False):
        model = BertQuantizedForSequenceClassification(cfg).eval()
        model.load_state_dict(sd, strict=True)
        model = model.cuda()
        model.bert.fused_attention = fused
        n0 = _attention_launches()
        with torch.no_grad():
            out = model(input_ids=ids, attention_mask=am, token_type_ids=tt, output_hidden_states=True)
        assert _attention_launches() - n0 == (2 if fused else 0)
        h = out.hidden_states[-1].cpu()
        err = (h - ref_h).abs()[valid]                    # hidden states of padding tokens are not consumed by anything
        spread = float(ref_h[valid].std())
        res[fused] = (float(err.mean()) / spread, float(err.max()) / spread)
        assert float(err.mean()) <= 0.03 * spread and float(err.max()) <= 0.75 * spread, (fused, res[fused])
        lerr = (out.logits.cpu() - ref_logits).abs()
        scale = max(float(ref_logits.abs().max()), 1e-3)
        assert float(lerr.max()) <= 0.1 * scale, (fused, float(lerr.max()), scale)
    # the fused kernel is no further from the reference than the op-by-op path (both see the same accumulation-order noise)
    assert res[True][0] <= 1.5 * res[False][0] + 1e-3, res
