"""
ORACLE — TEST INFRASTRUCTURE ONLY.

Functional restatement of the reference's quantized OPT forward
(models/opt_quantized/modeling_opt.py: attention :188-330, layer :360-441, decoder :784-853, CausalLM loss
:1084-1098) on top of oracle/oracle.py.  It consumes a plain state dict with HF OPT parameter names and the
per-layer quant config the reference's parser produces.  Used as (1) the CPU baseline `bench.py` times
("port" of the reference's CPU path — the reference itself cannot travel to the GPU box), and (2) an
on-device end-to-end comparator in tests (torch-CUDA emulation, same ops the reference would run on GPU).
Pinned against the reference's own forward on tests/golden/opt_tiny_*.npz (tests/test_oracle_models.py).
"""
import torch
import torch.nn.functional as F

from . import oracle as O


def _linear(x, sd, prefix, cfg, state):
    """PTQ linear: weights/bias quantised once (cached in `state`), input quantised every call (linear.py:59-76)."""
    if cfg.get("bypass", False):
        return F.linear(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"))
    if prefix not in state:
        w = O.operand_quantizer(cfg, "weight", False)(sd[prefix + ".weight"])
        b = sd.get(prefix + ".bias")
        if b is not None:
            b = O.operand_quantizer(cfg, "bias", False)(b)
        state[prefix] = (w, b)
    w, b = state[prefix]
    return O.gemm_linear(O.operand_quantizer(cfg, "data_in", True)(x), w, b)


def opt_layer_forward(h, sd, i, qc, num_heads, mask, state, do_layer_norm_before=True, act=F.relu):
    p = f"model.decoder.layers.{i}."
    lq = qc[f"model_layer_{i}"]
    bsz, tgt, H = h.shape
    d = H // num_heads
    residual = h
    if do_layer_norm_before:
        h = F.layer_norm(h, (H,), sd[p + "self_attn_layer_norm.weight"], sd[p + "self_attn_layer_norm.bias"])
    q = _linear(h, sd, p + "self_attn.q_proj", lq["self_attn"]["q_proj"], state) * (d ** -0.5)
    k = _linear(h, sd, p + "self_attn.k_proj", lq["self_attn"]["k_proj"], state)
    v = _linear(h, sd, p + "self_attn.v_proj", lq["self_attn"]["v_proj"], state)

    def shape(t):
        return t.view(bsz, tgt, num_heads, d).transpose(1, 2).contiguous().view(bsz * num_heads, tgt, d)

    q, k, v = shape(q), shape(k), shape(v)
    s = O.matmul_forward(q, k.transpose(1, 2), lq["self_attn"]["bmm_0"], style="bmm")
    s = s.view(bsz, num_heads, tgt, tgt) + mask
    s = torch.max(s, torch.tensor(torch.finfo(s.dtype).min, device=s.device)).view(bsz * num_heads, tgt, tgt)
    pr = F.softmax(s, dim=-1)
    o = O.matmul_forward(pr, v, lq["self_attn"]["bmm_1"], style="bmm")
    o = o.view(bsz, num_heads, tgt, d).transpose(1, 2).reshape(bsz, tgt, H)
    o = _linear(o, sd, p + "self_attn.out_proj", lq["self_attn"]["out_proj"], state)
    h = residual + o
    if not do_layer_norm_before:
        h = F.layer_norm(h, (H,), sd[p + "self_attn_layer_norm.weight"], sd[p + "self_attn_layer_norm.bias"])
    shp = h.shape
    h = h.reshape(-1, H)
    residual = h
    if do_layer_norm_before:
        h = F.layer_norm(h, (H,), sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"])
    h = _linear(h, sd, p + "fc1", lq["fc1"], state)
    h = act(h)
    h = _linear(h, sd, p + "fc2", lq["fc2"], state)
    h = (residual + h).view(shp)
    if not do_layer_norm_before:
        h = F.layer_norm(h, (H,), sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"])
    return h


def causal_mask(bsz, tgt, dtype, device):
    neg = torch.finfo(dtype).min
    m = torch.triu(torch.full((tgt, tgt), neg, dtype=dtype, device=device), diagonal=1)
    return m[None, None].expand(bsz, 1, tgt, tgt)


def opt_forward(sd, qc, input_ids, num_layers, num_heads, labels=None, state=None, collect=None):
    """Full forward: returns (logits, loss).  `state` caches the PTQ-quantised weights across calls; `collect` (a list)
    receives the INPUT hidden state of every decoder layer (the reference's output_hidden_states[:-1])."""
    state = {} if state is None else state
    bsz, tgt = input_ids.shape
    emb = sd["model.decoder.embed_tokens.weight"]
    pos = sd["model.decoder.embed_positions.weight"]
    h = F.embedding(input_ids, emb) + pos[torch.arange(tgt, device=input_ids.device) + 2][None]
    mask = causal_mask(bsz, tgt, h.dtype, h.device)
    for i in range(num_layers):
        if collect is not None:
            collect.append(h)
        h = opt_layer_forward(h, sd, i, qc, num_heads, mask, state)
    H = h.shape[-1]
    if "model.decoder.final_layer_norm.weight" in sd:
        h = F.layer_norm(h, (H,), sd["model.decoder.final_layer_norm.weight"], sd["model.decoder.final_layer_norm.bias"])
    logits = O.gemm_linear(h, sd.get("lm_head.weight", emb))
    loss = None
    if labels is not None:
        loss = F.cross_entropy(logits[:, :-1].reshape(-1, logits.shape[-1]), labels[:, 1:].reshape(-1))
    return logits, loss
