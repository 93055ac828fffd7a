"""Per-op comparison of one OPT decoder layer: our modules (op by op and fused) vs the oracle on identical inputs."""
import os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle import opt_ref, oracle as O
import test_gpu_fused_glue as T

def stat(name, a, b):
    e = (a - b).abs()
    print(f"{name:28s} mean|err| {float(e.mean()):.3e}  max {float(e.max()):.3e}  ref std {float(b.std()):.3e}  frac>1e-3std {float((e > 1e-3 * b.std()).float().mean()):.2e}")

for width in (6, 4):
    print("=== width", width)
    model = T._opt_model(width=width)
    dec = model.model.decoder
    ids = torch.randint(0, 512, (3, 128), device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    with torch.no_grad():
        dec.fused_glue, dec.fused_attention = False, False
        model(input_ids=ids)   # PTQ
        sd = {k: v.detach() for k, v in model.state_dict().items()}
        qc = model.config.quant_config
        lq = qc["model_layer_0"]
        L0 = dec.layers[0]
        at = L0.self_attn
        h = dec.embed_tokens(ids) + dec.embed_positions(torch.ones_like(ids), 0)
        x = L0.self_attn_layer_norm(h)
        state = {}
        p = "model.decoder.layers.0."
        for nm, mod, cfg in (("q_proj", at.q_proj, lq["self_attn"]["q_proj"]), ("k_proj", at.k_proj, lq["self_attn"]["k_proj"]), ("v_proj", at.v_proj, lq["self_attn"]["v_proj"])):
            ours = mod(x)
            ref = opt_ref._linear(x, sd, p + "self_attn." + nm, cfg, state)
            stat(nm, ours, ref)
        mask = opt_ref.causal_mask(3, 128, torch.float32, "cuda")
        ref_layer = opt_ref.opt_layer_forward(h, sd, 0, qc, 4, mask, {})
        cm = None
        from llm_mixed_q_b200.models.opt_quantized.modeling_opt import _causal_additive_mask
        cm = _causal_additive_mask(torch.ones(3, 128, dtype=torch.bool, device="cuda"), 3, 128, torch.float32, "cuda")
        o_unf, _ = L0(h, attention_mask=cm, causal_only=False, fused_glue=False)
        o_att, _ = L0(h, attention_mask=cm, causal_only=True, fused_glue=False)
        o_fus, _ = L0(h, attention_mask=cm, causal_only=True, fused_glue=True)
        stat("layer op-by-op vs oracle", o_unf, ref_layer)
        stat("layer fused-attn vs oracle", o_att, ref_layer)
        stat("layer fused-glue vs oracle", o_fus, ref_layer)
        stat("layer fused-glue vs op-by-op", o_fus, o_unf)
        # attention piece only
        q = at.q_proj(x) * at.scaling; k = at.k_proj(x); v = at.v_proj(x)
        sys.path.insert(0, "tests")
        import test_gpu_consumers as TC
        cfg = lq["self_attn"]["bmm_0"]
        ref_a, _ = TC._oracle_attention(q, k, v, cfg, 4)
        from llm_mixed_q_b200.models.quantize.quantized_functions.attention import fused_causal_attention
        stat("attention fused vs oracle", fused_causal_attention(q, k, v, cfg, cfg, 4), ref_a)
        # fc path
        h2 = torch.randn(3 * 128, 256, device="cuda")
        x2 = L0.final_layer_norm(h2)
        f1 = L0.fc1(x2); r1 = opt_ref._linear(x2, sd, p + "fc1", lq["fc1"], state); stat("fc1", f1, r1)
        a = torch.relu(r1)
        f2 = L0.fc2(a); r2 = opt_ref._linear(a, sd, p + "fc2", lq["fc2"], state); stat("fc2", f2, r2)
