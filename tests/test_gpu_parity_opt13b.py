"""GPU: FALSIFIABLE end-to-end parity at the headline model (BASELINE configs[2]'s OPT-1.3B, W6A6 block_fp, 1 x 2048 tokens).

Anchors (tests/golden/opt13b_bfp6.npz, written by oracle/gen_golden_opt13b.py from the UNMODIFIED reference's CPU forward):
loss, per-token log-partition, a 128 x 786 logits sample, and a 64 x 128 slice of every decoder layer's input hidden state.

Rounding makes the quantised forward discontinuous: two CORRECT implementations that differ by an ulp somewhere (LayerNorm
statistics, exp, a division) drift apart by one-step rounding flips that diffuse through 24 layers, so logits cannot be compared
bit for bit.  Instead of asserting a loose tolerance, the tests MEASURE that drift with controls and hold the CUDA path to it:

  control A  the oracle port run on the same B200 with torch-CUDA fp32 ops against the reference's CPU golden — the reference
             against itself on two back ends (measured: mean |dlogit| = 5.4 % of the logit spread at 24 layers);
  control B  the same oracle with every contraction run in reversed K order (oracle.ACCUMULATION_ORDER) against itself
             (measured: 8e-7 — products of block-quantised operands are exact and their fp32 sums nearly so; GEMM accumulation
             order is NOT what seeds the flips);
  control C  per layer, teacher-forced: the oracle on the host cores against the oracle on the GPU on the same layer input.

1. full forward: mean |dlogit| of the fused CUDA path against the golden must be within 1.25x of control A's, and the hidden
   state after every layer within 1.5x of control A's distance from the golden at that layer.
2. teacher-forced per layer: every fused decoder layer is fed the oracle trajectory's input of that layer (itself checked against
   the golden slices) and its output compared with the oracle layer's output on the SAME input — no compounding across layers, so
   a wrong scale / mask / block orientation in any one layer shows up as >= 10 % of the layer's update against a flip floor
   below 1 %; held to 2x control C and to the one-step-flip model on the first-order quantised tensor.
"""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLD

pytestmark = pytest.mark.gpu

OPT13B = dict(hidden_size=2048, num_hidden_layers=24, ffn_dim=8192, num_attention_heads=32, vocab_size=50272,
              max_position_embeddings=2048)
L, HEADS = 24, 32


def _rms(t):
    return float(t.double().pow(2).mean().sqrt())


@pytest.fixture(scope="module")
def setup():
    from llm_mixed_q_b200.models.opt_quantized import OPTQuantizedConfig, OPTQuantizedForCausalLM, parse_opt_quantized_config

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    z = np.load(os.path.join(GOLD, "opt13b_bfp6.npz"))
    raw = json.load(open(os.path.join(GOLD, "configs.json")))["raw"]["bfp_6bit.toml"]
    torch.manual_seed(0)
    model = OPTQuantizedForCausalLM(OPTQuantizedConfig(quant_config=json.loads(json.dumps(raw)), tie_word_embeddings=False,
                                                       **OPT13B)).eval()
    sd = model.state_dict()
    for k, (s, a) in zip(z["checksum_keys"], z["checksum_vals"]):
        v = sd[str(k)].double()
        assert abs(float(v.sum()) - s) <= 1e-9 * (a + 1e-30) and abs(float(v.abs().sum()) - a) <= 1e-9 * a, \
            f"seeded init of {k} differs from the reference's"
    sd0 = {k: v.detach().clone().cuda() for k, v in sd.items()}          # before the PTQ overwrite
    model = model.cuda()
    qc = parse_opt_quantized_config(json.loads(json.dumps(raw)), L)
    ids = torch.from_numpy(z["input_ids"]).cuda()
    return model, sd0, qc, ids, z


def _oracle_forward(sd0, qc, ids, order):
    from oracle import opt_ref, oracle as O

    O.ACCUMULATION_ORDER = order
    try:
        hs = []
        with torch.no_grad():
            logits, loss = opt_ref.opt_forward(sd0, qc, ids, L, HEADS, labels=ids, collect=hs)
    finally:
        O.ACCUMULATION_ORDER = "natural"
    return logits[0], float(loss), hs


def _against_golden(logits, loss, z):
    spread = float(z["logits_std"])
    err = (logits[::16, ::64].cpu() - torch.from_numpy(z["logits_sub"])).abs()
    lse = torch.logsumexp(logits.double(), -1).cpu()
    return dict(loss=loss, dloss_rel=abs(loss - float(z["loss"])) / float(z["loss"]), mean=float(err.mean()) / spread,
                max=float(err.max()) / spread, dlse=float((lse - torch.from_numpy(z["logits_row_lse"])).abs().max()))


@pytest.mark.timeout(1200)
def test_opt13b_full_forward_is_within_the_measured_noise_floor(setup):
    model, sd0, qc, ids, z = setup
    ts, ds = (int(v) for v in z["sub_strides"])
    gold_h = torch.from_numpy(z["h_in_sub"])
    upd = torch.from_numpy(np.concatenate([z["update_rms"], z["update_rms"][-1:]]))

    # control A: oracle port on this GPU (cuBLAS fp32) vs the reference's CPU forward
    lg_a, loss_a, hs_a = _oracle_forward(sd0, qc, ids, "natural")
    ctl_a = _against_golden(lg_a, loss_a, z)
    h_err_a = [_rms(hs_a[i][0, ::ts, ::ds].cpu() - gold_h[i]) / float(upd[max(i - 1, 0)]) for i in range(L)]
    # control B: same oracle, reversed accumulation order, vs itself
    lg_b, loss_b, hs_b = _oracle_forward(sd0, qc, ids, "reversed")
    spread = float(z["logits_std"])
    ctl_b = dict(mean=float((lg_b - lg_a).abs().mean()) / spread, max=float((lg_b - lg_a).abs().max()) / spread,
                 dloss_rel=abs(loss_b - loss_a) / loss_a)
    del lg_b, hs_b

    # the CUDA path: fused layers (and once op by op)
    dec = model.model.decoder
    res, h_err = {}, {}
    for name, fused in (("fused", True), ("op_by_op", False)):
        for k, v in model.state_dict().items():
            v.copy_(sd0[k])
        for m in model.modules():
            if hasattr(m, "weight_requires_quantisation"):
                m.weight_requires_quantisation = True
        dec.fused_glue, dec.fused_attention = fused, fused
        if fused:
            assert dec.layers[0]._fused_plan(ids.shape[1]) is not None
        with torch.no_grad():
            out = model(input_ids=ids, labels=ids, output_hidden_states=True)
        res[name] = _against_golden(out.logits[0], float(out.loss), z)
        res[name]["vs_oracle_gpu_mean"] = float((out.logits[0] - lg_a).abs().mean()) / spread
        h_err[name] = [_rms(out.hidden_states[i][0, ::ts, ::ds].cpu() - gold_h[i]) / float(upd[max(i - 1, 0)]) for i in range(L)]
        del out
    dec.fused_glue, dec.fused_attention = True, True
    report = dict(control_A_oracle_gpu_vs_reference_cpu=ctl_a, control_B_reversed_K_vs_natural=ctl_b, cuda_path=res,
                  hidden_err_over_update_rms=dict(control_A=h_err_a[::4] + h_err_a[-1:], fused=h_err["fused"][::4] + h_err["fused"][-1:],
                                                  op_by_op=h_err["op_by_op"][::4] + h_err["op_by_op"][-1:]))
    print("OPT13B_PARITY " + json.dumps(report))
    os.makedirs(os.path.join(os.path.dirname(GOLD), "..", "gpurun_out"), exist_ok=True)
    with open(os.path.join(os.path.dirname(GOLD), "..", "gpurun_out", "parity_opt13b_full.json"), "w") as f:
        json.dump(report, f, indent=1)

    floor = ctl_a["mean"]
    assert ctl_a["dloss_rel"] <= 2e-4, ctl_a                                   # the oracle itself reproduces the reference's loss
    for name, r in res.items():
        assert r["dloss_rel"] <= 2e-4, (name, r)
        assert r["dlse"] <= max(1.5 * ctl_a["dlse"], 2e-3), (name, r, ctl_a)
        assert r["mean"] <= 1.25 * floor + 1e-3, (name, r["mean"], floor)
        assert r["max"] <= max(1.5 * ctl_a["max"], 0.1), (name, r["max"], ctl_a["max"])
        # hidden states along the way: not further from the reference than the control at ANY layer (x1.5 + an absolute floor)
        for i in range(L):
            assert h_err[name][i] <= 1.5 * h_err_a[i] + 2e-3, (name, i, h_err[name][i], h_err_a[i])


CPU_CONTROL_LAYERS = (0, 1, 9, 16, 23)          # layers on which the oracle is ALSO run on the host cores (a few seconds each)


@pytest.mark.timeout(1500)
def test_opt13b_teacher_forced_layers_match_the_oracle_layer_by_layer(setup):
    """Every fused decoder layer on the reference trajectory's input of that layer, against the oracle layer on the SAME input.
    Reversing the contraction order changes almost nothing (control B is ~1e-9: products of block-quantised operands are exact
    and their fp32 sums nearly so); what seeds rounding flips is the ulp-level difference between two implementations of
    LayerNorm statistics / exp / division — exactly what separates torch-CPU from torch-CUDA.  So the per-layer control is the
    oracle on the host cores against the oracle on the GPU (the reference against itself on two back ends), on a subset of layers;
    the fused layer must stay within 2x of it, and under an absolute 2 % of the layer's update everywhere (a wrong scale / mask /
    block orientation / missing bias gives >= 10 %).  First-order check of the one-step-flip model: LayerNorm + x-quantizer
    output (identical input) — every mismatching element is off by exactly ONE quantisation step, and <= 2e-3 of them are."""
    from llm_mixed_q_b200.models.quantize.quantized_functions.fused_glue import norm_quantize
    from oracle import opt_ref, oracle as O

    model, sd0, qc, ids, z = setup
    ts, ds = (int(v) for v in z["sub_strides"])
    gold_h = torch.from_numpy(z["h_in_sub"])
    dec = model.model.decoder
    for k, v in model.state_dict().items():
        v.copy_(sd0[k])
    for m in model.modules():
        if hasattr(m, "weight_requires_quantisation"):
            m.weight_requires_quantisation = True
    dec.fused_glue, dec.fused_attention = True, True
    S = ids.shape[1]
    mask = opt_ref.causal_mask(1, S, torch.float32, ids.device)
    mask_cpu = opt_ref.causal_mask(1, S, torch.float32, "cpu")
    state_n, state_r = {}, {}
    emb, pos = sd0["model.decoder.embed_tokens.weight"], sd0["model.decoder.embed_positions.weight"]
    h = torch.nn.functional.embedding(ids, emb) + pos[torch.arange(S, device=ids.device) + 2][None]
    torch.set_num_threads(os.cpu_count() or 1)
    rows = []
    with torch.no_grad():
        for i in range(L):
            # the teacher trajectory is the reference's: its sampled slice must sit on the golden within the drift of control A
            traj = _rms(h[0, ::ts, ::ds].cpu() - gold_h[i])
            ref = opt_ref.opt_layer_forward(h, sd0, i, qc, HEADS, mask, state_n)
            O.ACCUMULATION_ORDER = "reversed"
            try:
                ctl = opt_ref.opt_layer_forward(h, sd0, i, qc, HEADS, mask, state_r)
            finally:
                O.ACCUMULATION_ORDER = "natural"
            state_r.clear()
            lyr = dec.layers[i]
            plan = lyr._fused_plan(S)
            assert plan is not None
            ours, _ = lyr(h, attention_mask=None, causal_only=True, fused_glue=True)
            u = _rms(ref - h)
            d_o, d_c = ours - ref, ctl - ref
            big = 0.05 * u
            row = dict(layer=i, traj_err=traj, update_rms=u, ours=_rms(d_o) / u, control_reversed_k=_rms(d_c) / u,
                       ours_max=float(d_o.abs().max()) / u, ours_frac_big=float((d_o.abs() > big).float().mean()))
            # first-order tensor: LayerNorm + q_proj's x-quantizer on the identical input
            ln = lyr.self_attn_layer_norm
            (xq,) = norm_quantize(h, ln.weight, ln.bias, ln.eps, [plan["q_in"]])
            pfx = f"model.decoder.layers.{i}.self_attn_layer_norm."
            xo = O.operand_quantizer(qc[f"model_layer_{i}"]["self_attn"]["q_proj"], "data_in", True)(
                torch.nn.functional.layer_norm(h, (h.shape[-1],), sd0[pfx + "weight"], sd0[pfx + "bias"]))
            dq = (xq.float() - xo).abs().view(-1, 16)
            step_hi = xo.abs().view(-1, 16).amax(1, keepdim=True) / 16.0          # one step = 2^(e-5) in (max/31, max/16]
            mism = dq > 0
            row["lnq_mismatch_frac"] = float(mism.float().mean())
            row["lnq_more_than_one_step"] = int((dq > step_hi * (1 + 1e-6)).sum())
            if i in CPU_CONTROL_LAYERS:
                p = f"model.decoder.layers.{i}."
                sd_cpu = {k: v.cpu() for k, v in sd0.items() if k.startswith(p)}
                ref_cpu = opt_ref.opt_layer_forward(h.cpu(), sd_cpu, i, qc, HEADS, mask_cpu, {})
                d_h = ref_cpu.to(ref.device) - ref
                row["control_cpu_vs_gpu"] = _rms(d_h) / u
                row["control_cpu_vs_gpu_frac_big"] = float((d_h.abs() > big).float().mean())
                row["ours_vs_cpu"] = _rms(ours - ref_cpu.to(ref.device)) / u
            rows.append(row)
            h = ref
    print("OPT13B_TEACHER_FORCED " + json.dumps(rows))
    with open(os.path.join(os.path.dirname(GOLD), "..", "gpurun_out", "parity_opt13b_layers.json"), "w") as f:
        json.dump(rows, f, indent=1)
    ctl_mean = sum(r["control_cpu_vs_gpu"] for r in rows if "control_cpu_vs_gpu" in r) / len(CPU_CONTROL_LAYERS)
    ours_mean = sum(r["ours"] for r in rows) / len(rows)
    for r in rows:
        assert r["ours"] <= 0.02, r                                            # absolute: 2 % of the layer's update
        assert r["lnq_more_than_one_step"] == 0 and r["lnq_mismatch_frac"] <= 2e-3, r
        if "control_cpu_vs_gpu" in r:
            assert r["ours"] <= 2.0 * r["control_cpu_vs_gpu"] + 2e-3, r
    assert ours_mean <= 1.5 * ctl_mean + 1e-3, (ours_mean, ctl_mean)
