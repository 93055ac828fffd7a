import json, math, os, sys, torch
sys.path.insert(0, os.getcwd())
from llm_mixed_q_b200.models.llama_quantized import LlamaQuantizedConfig, LlamaQuantizedForCausalLM
from llm_mixed_q_b200.models.quantize import get_quantized_func
from llm_mixed_q_b200.models.quantize.quantized_functions.split_attention import rope_quantize_split, split_attention
from llm_mixed_q_b200.models.quantize.quantized_functions.fused_glue import norm_quantize
which = sys.argv[1]
if which == "w8":
    qc = json.load(open("tests/golden/configs.json"))["raw"]["block_log.toml"]
else:
    qc = "configs/llama_w4a4_block_log.toml"
cfg = LlamaQuantizedConfig(quant_config=qc, num_hidden_layers=1, vocab_size=1024)
torch.manual_seed(0)
with torch.device("cuda"):
    model = LlamaQuantizedForCausalLM(cfg).eval()
layer = model.model.layers[0]; at = layer.self_attn; qcfg = at.quant_config
g = torch.Generator(device="cuda").manual_seed(1)
h = torch.randn(1, 2048, 4096, device="cuda", generator=g)
B, S, H = h.shape; nh, d = at.num_heads, at.head_dim
pos = torch.arange(S, device="cuda")[None]
rms = lambda t: float(t.double().pow(2).mean().sqrt())
rel = lambda a, b: rms(a.float() - b.float()) / (rms(b.float()) + 1e-30)
with torch.no_grad():
    plan = layer._fused_plan(S)
    # --- op by op pieces
    x = layer.input_layernorm(h)
    q_o, k_o, v_o = at.q_proj(x), at.k_proj(x), at.v_proj(x)
    # --- fused pieces
    xq, xk, xv = norm_quantize(h, layer.input_layernorm.weight, None, layer.input_layernorm.variance_epsilon, [plan["q_in"], plan["k_in"], plan["v_in"]])
    q_f, k_f, v_f = at.q_proj.forward_prequantized(xq), at.k_proj.forward_prequantized(xk), at.v_proj.forward_prequantized(xv)
    print("q", rel(q_f.view_as(q_o), q_o), "k", rel(k_f.view_as(k_o), k_o), "v", rel(v_f.view_as(v_o), v_o))
    # attention op by op on the op-by-op q,k,v
    shp = (B, S, nh, d)
    qs, ks, vs = (t.view(*shp).transpose(1, 2) for t in (q_o, k_o, v_o))
    cos, sin = at.rotary_emb(vs, seq_len=S)
    rope_cfg = qcfg["rotary_positional_encoding"]
    qr, kr = get_quantized_func("rotary_positional_encoding", rope_cfg)(qs, ks, cos, sin, pos, rope_cfg)
    mm0 = get_quantized_func("matmul", qcfg["matmul_0"])
    sc = mm0(qr, kr.transpose(2, 3), config=qcfg["matmul_0"]) / math.sqrt(d)
    mask = torch.triu(torch.full((S, S), torch.finfo(torch.float32).min, device="cuda"), diagonal=1)[None, None]
    sc2 = torch.max(sc + mask, torch.tensor(torch.finfo(torch.float32).min, device="cuda"))
    p = torch.softmax(sc2, dim=-1, dtype=torch.float32)
    mm1 = get_quantized_func("matmul", qcfg["matmul_1"])
    o_o = mm1(p, vs, config=qcfg["matmul_1"]).transpose(1, 2).reshape(B, S, H)
    # split attention on the SAME op-by-op q,k,v
    Qq, Kp = rope_quantize_split(q_o.view(B, S, H), k_o.view(B, S, H), cos, sin, None, rope_cfg, qcfg["matmul_0"], nh)
    o_s = split_attention(Qq, Kp, v_o.view(B, S, H), qcfg["matmul_1"], nh, math.sqrt(d), causal=True)
    print("attention out (same q,k,v):", rel(o_s, o_o), "max", float((o_s - o_o).abs().max()), "rms ref", rms(o_o))
    # scores check
    from llm_mixed_q_b200 import _lib as L
    import ctypes
    lib = L.load()
    scores = torch.empty((nh, S, S), device="cuda")
    TA = (ctypes.c_int32 * 3)(0, 0, 0); TB = (ctypes.c_int32 * 3)(2, 1, 0)
    L.check(lib.bq_bmm_split_tn(Qq[0].data_ptr(), Kp[0].data_ptr(), scores.data_ptr(), nh, S, S, d, 1, 3, 3, TA, TB, S, S * S, 0, L.stream_ptr(h.device)), "x")
    tri = torch.tril(torch.ones(S, S, device="cuda", dtype=torch.bool))
    print("scores (lower tri):", rel((scores / math.sqrt(d))[:, tri], sc[0][:, tri]))
    # P check: quantised P from op-by-op
    from oracle import oracle as O
    kind, kw, bs = ("block_log", None, None)
    w = qcfg["matmul_1"]["data_in_width"]; ebw = qcfg["matmul_1"]["data_in_exponent_bias_width"]
    pq = O.block_log_quantize(p.reshape(nh, S, S), w, ebw, [1, 16], True)
    print("width", w, "ebw", ebw, "P upper-tri fill value (reference):", float(pq[0, 0, 100]), "sum of fill row 0:", float(pq[0, 0, 1:].sum()), "p00", float(pq[0,0,0]))
    # ---- later stages
    from llm_mixed_q_b200.models.quantize.quantized_modules.linear import quantize_operand_bf16
    from llm_mixed_q_b200.models.quantize.quantized_functions.fused_glue import silu_mul_quantize
    h2_o = h + at.o_proj(o_o)
    okind, okw = plan["o_in"]
    oq = quantize_operand_bf16(o_o.reshape(B * S, H), okind, okw, [1, 16], True)
    h2_f = at.o_proj.forward_prequantized(oq, residual=h).view(B, S, H)
    print("h2 update (same attention out):", rel(h2_f - h, h2_o - h))
    n2 = layer.post_attention_layernorm; mlp = layer.mlp
    x2 = n2(h2_o)
    g_o, u_o = mlp.gate_proj(x2), mlp.up_proj(x2)
    a_o = mlp.act_fn(g_o) * u_o
    d_o = mlp.down_proj(a_o)
    xg, xu = norm_quantize(h2_o, n2.weight, None, n2.variance_epsilon, [plan["gate_in"], plan["up_in"]])
    g_f, u_f = mlp.gate_proj.forward_prequantized(xg), mlp.up_proj.forward_prequantized(xu)
    print("gate", rel(g_f.view_as(g_o), g_o), "up", rel(u_f.view_as(u_o), u_o))
    a_f = silu_mul_quantize(g_o.view(B * S, -1), u_o.view(B * S, -1), plan["down_in"])
    d_f = mlp.down_proj.forward_prequantized(a_f).view(B, S, H)
    print("down (same gate/up):", rel(d_f, d_o), "rms", rms(d_o))
    a_f2 = silu_mul_quantize(g_f.view(B * S, -1), u_f.view(B * S, -1), plan["down_in"])
    d_f2 = mlp.down_proj.forward_prequantized(a_f2).view(B, S, H)
    print("down (fused gate/up):", rel(d_f2, d_o))
    # the down_proj x operand itself
    dk, dkw = plan["down_in"]
    aq_o = quantize_operand_bf16(a_o.reshape(B * S, -1), dk, dkw, [1, 16], True)
    print("down_proj x operand: fused kernel vs quantizer(op-by-op product):", rel(a_f, aq_o), "mismatch frac", float((a_f != aq_o).float().mean()))
