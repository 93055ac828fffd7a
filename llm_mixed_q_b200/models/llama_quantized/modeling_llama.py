"""
Quantized Llama — HF-compatible parameter names (`model.layers.<i>.self_attn.q_proj.weight`, …).

Quantisation wiring of reference llama_quantized/modeling_llama.py: seven quantized Linears without bias
(:208-210, :237-240), RoPE with quantised cos/sin tables (:289-299), two quantized 4-D matmuls
(:309-314, :341-344; scores divided by sqrt(d) AFTER matmul_0), fp32 RMSNorm / softmax / SiLU / lm_head (:772).
Forward-only perplexity path; no KV cache.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn
from torch.nn import CrossEntropyLoss
from transformers.activations import ACT2FN
from transformers.modeling_outputs import CausalLMOutputWithPast, SequenceClassifierOutputWithPast
from transformers.modeling_utils import PreTrainedModel

from ..quantize import get_quantized_cls, get_quantized_func
from ..quantize.quantized_functions.attention import (causal_key_mask, fusable as _attn_fusable, fused_causal_attention_q,
                                                       output_quantizable, quantize_qkv)
from ..quantize.quantized_functions.fp32_linear import fp32_linear
from ..quantize.quantized_functions.loss import causal_lm_loss
from ..quantize.quantized_functions.fused_glue import (_NORM_KINDS, linear_input_format, norm_quantize, row_block16_format,
                                                        silu_mul_quantize)
from ..quantize.quantized_functions.split_attention import rope_quantize_split, split_attention, splittable as _attn_splittable
from ..quantize.quantized_functions.rotary_positional_encoding import apply_token_major as _rope_token_major
from ..quantize.quantized_functions.rotary_positional_encoding import apply_token_major_quantized as _rope_token_major_quantized
from ..quantize.quantized_functions.rotary_positional_encoding import rope_quantize_operands
from ..quantize.quantized_modules.linear import (gated_silu_fusable, gated_silu_prequantized, operand_format, qkv_plain_fusable,
                                                 qkv_plain_prequantized, qkv_rope_fusable, qkv_rope_prequantized, quantize_operand_bf16, rope_epilogue_fusable, rope_prequantized)
from .configuration_llama import LlamaQuantizedConfig


class LlamaRMSNorm(nn.Module):
    def __init__(self, hidden_size, eps=1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(hidden_size))
        self.variance_epsilon = eps

    def forward(self, hidden_states):
        dtype = hidden_states.dtype
        variance = hidden_states.to(torch.float32).pow(2).mean(-1, keepdim=True)
        hidden_states = hidden_states * torch.rsqrt(variance + self.variance_epsilon)
        return (self.weight * hidden_states).to(dtype)


class LlamaRotaryEmbedding(nn.Module):
    """cos/sin tables [1, 1, seq, dim] (reference modeling_llama.py:119-165)."""

    def __init__(self, dim, max_position_embeddings=2048, base=10000, device=None):
        super().__init__()
        self.dim, self.base = dim, base
        inv_freq = 1.0 / (base ** (torch.arange(0, dim, 2).float().to(device) / dim))
        self.register_buffer("inv_freq", inv_freq, persistent=False)
        self._build(max_position_embeddings, device=inv_freq.device)

    def _build(self, seq_len, device):
        self.max_seq_len_cached = seq_len
        t = torch.arange(seq_len, device=device, dtype=self.inv_freq.dtype)
        freqs = torch.einsum("i,j->ij", t, self.inv_freq.to(device))
        emb = torch.cat((freqs, freqs), dim=-1)
        self.register_buffer("cos_cached", emb.cos()[None, None, :, :], persistent=False)
        self.register_buffer("sin_cached", emb.sin()[None, None, :, :], persistent=False)

    def forward(self, x, seq_len=None):
        if seq_len > self.max_seq_len_cached or self.cos_cached.device != x.device:
            self._build(max(seq_len, self.max_seq_len_cached), device=x.device)
        return self.cos_cached[:, :, :seq_len, ...].to(dtype=x.dtype), self.sin_cached[:, :, :seq_len, ...].to(dtype=x.dtype)


class LlamaQuantizedMLP(nn.Module):
    def __init__(self, hidden_size: int, intermediate_size: int, hidden_act: str, quant_config: dict):
        super().__init__()
        qc = quant_config
        self.gate_proj = get_quantized_cls("linear", qc["gate_proj"])(hidden_size, intermediate_size, bias=False, config=qc["gate_proj"])
        self.down_proj = get_quantized_cls("linear", qc["down_proj"])(intermediate_size, hidden_size, bias=False, config=qc["down_proj"])
        self.up_proj = get_quantized_cls("linear", qc["up_proj"])(hidden_size, intermediate_size, bias=False, config=qc["up_proj"])
        self.act_fn = ACT2FN[hidden_act]
        self.hidden_act = hidden_act
        self.quant_config = qc

    def forward(self, x):
        return self.down_proj(self.act_fn(self.gate_proj(x)) * self.up_proj(x))


class LlamaQuantizedAttention(nn.Module):
    def __init__(self, config: LlamaQuantizedConfig, layer_id: int = 0):
        super().__init__()
        self.config = config
        self.hidden_size = config.hidden_size
        self.num_heads = config.num_attention_heads
        self.head_dim = self.hidden_size // self.num_heads
        self.max_position_embeddings = config.max_position_embeddings
        if self.head_dim * self.num_heads != self.hidden_size:
            raise ValueError(f"hidden_size must be divisible by num_heads (got `hidden_size`: {self.hidden_size} and `num_heads`: {self.num_heads}).")
        qc = config.quant_config[f"model_layer_{layer_id}"]["self_attn"]
        H = self.num_heads * self.head_dim
        self.q_proj = get_quantized_cls("linear", qc["q_proj"])(self.hidden_size, H, bias=False, config=qc["q_proj"])
        self.k_proj = get_quantized_cls("linear", qc["k_proj"])(self.hidden_size, H, bias=False, config=qc["k_proj"])
        self.v_proj = get_quantized_cls("linear", qc["v_proj"])(self.hidden_size, H, bias=False, config=qc["v_proj"])
        self.o_proj = get_quantized_cls("linear", qc["o_proj"])(H, self.hidden_size, bias=False, config=qc["o_proj"])
        self.rotary_emb = LlamaRotaryEmbedding(self.head_dim, max_position_embeddings=self.max_position_embeddings)
        self.quant_config = qc

    def forward(self, hidden_states, attention_mask=None, position_ids=None, output_attentions=False):
        bsz, q_len, _ = hidden_states.size()
        shp = (bsz, q_len, self.num_heads, self.head_dim)
        query_states = self.q_proj(hidden_states).view(*shp).transpose(1, 2)
        key_states = self.k_proj(hidden_states).view(*shp).transpose(1, 2)
        value_states = self.v_proj(hidden_states).view(*shp).transpose(1, 2)

        cos, sin = self.rotary_emb(value_states, seq_len=q_len)
        rope_cfg = self.quant_config["rotary_positional_encoding"]
        query_states, key_states = get_quantized_func("rotary_positional_encoding", rope_cfg)(
            query_states, key_states, cos, sin, position_ids, rope_cfg)

        mm0 = get_quantized_func("matmul", self.quant_config["matmul_0"])
        attn_weights = mm0(query_states, key_states.transpose(2, 3), config=self.quant_config["matmul_0"]) / math.sqrt(self.head_dim)
        if attention_mask is not None:
            attn_weights = attn_weights + attention_mask
            attn_weights = torch.max(attn_weights, torch.tensor(torch.finfo(attn_weights.dtype).min, device=attn_weights.device))
        attn_weights = nn.functional.softmax(attn_weights, dim=-1, dtype=torch.float32).to(query_states.dtype)

        mm1 = get_quantized_func("matmul", self.quant_config["matmul_1"])
        attn_output = mm1(attn_weights, value_states, config=self.quant_config["matmul_1"])
        attn_output = attn_output.transpose(1, 2).reshape(bsz, q_len, self.hidden_size)
        attn_output = self.o_proj(attn_output)
        return attn_output, (attn_weights if output_attentions else None)


class LlamaQuantizedDecoderLayer(nn.Module):
    def __init__(self, config: LlamaQuantizedConfig, layer_id: int = 0):
        super().__init__()
        self.hidden_size = config.hidden_size
        self.self_attn = LlamaQuantizedAttention(config=config, layer_id=layer_id)
        self.mlp = LlamaQuantizedMLP(self.hidden_size, config.intermediate_size, config.hidden_act,
                                     config.quant_config[f"model_layer_{layer_id}"]["mlp"])
        self.input_layernorm = LlamaRMSNorm(config.hidden_size, eps=config.rms_norm_eps)
        self.post_attention_layernorm = LlamaRMSNorm(config.hidden_size, eps=config.rms_norm_eps)
        self.fused_rope = True           # False: RoPE with torch ops + separate q / k^T quantizer launches (A/B, tests)

    # ------------------------------------------------------------------ fused layer (PTQ inference, causal mask)
    def _fused_plan(self, seq_len: int):
        """Formats for running the layer on fused kernels (same idea as OPTQuantizedDecoderLayer._fused_plan): RMSNorm+quantize,
        GEMMs that read bf16 operands and apply the next op's x-quantizer / the residual add in their epilogue, one attention
        kernel.  None when any piece is not eligible (block_log, odd block shapes, > 8 significant bits, QAT...)."""
        cache = getattr(self, "_plan_cache", None)
        if cache is not None and cache[0] == seq_len:
            return cache[1]
        at, mlp = self.self_attn, self.mlp
        H, d = self.hidden_size, at.head_dim
        plan = None
        qc = at.quant_config
        if (seq_len % 16 == 0 and H % 32 == 0 and mlp.gate_proj.out_features % 16 == 0 and mlp.hidden_act == "silu"
                and _attn_fusable(qc["matmul_0"], qc["matmul_1"], d, seq_len) and output_quantizable(at.o_proj.config, H, seq_len)):
            # every Llama Linear sees the 3-D [B, S, H] activation (reference modeling_llama.py:274-276, :347, :422)
            lin = lambda m: linear_input_format(m, rows=seq_len)
            fmts = dict(q_in=lin(at.q_proj), k_in=lin(at.k_proj), v_in=lin(at.v_proj), o_in=lin(at.o_proj),
                        gate_in=lin(mlp.gate_proj), up_in=lin(mlp.up_proj), down_in=lin(mlp.down_proj),
                        v_out=row_block16_format(qc["matmul_1"], "weight", d, rows=seq_len))
            if all(v is not None for v in fmts.values()):
                plan = fmts
        if (plan is None and seq_len % 16 == 0 and H % 32 == 0 and mlp.gate_proj.out_features % 16 == 0 and mlp.hidden_act == "silu"
                and _attn_splittable(qc["matmul_0"], qc["matmul_1"], d, seq_len)):
            # block_log: matmul_0 / matmul_1 keep k / v in fp32 (reference matmul.py:286-297) -> the three-kernel attention of
            # split_attention.py; the Linears' x-quantizers still run inside RMSNorm / SiLU*up (block-local carrier rule, bq.h)
            lin = lambda m: linear_input_format(m, rows=seq_len, kinds=_NORM_KINDS)
            fmts = dict(mode="split", q_in=lin(at.q_proj), k_in=lin(at.k_proj), v_in=lin(at.v_proj), o_in=lin(at.o_proj),
                        gate_in=lin(mlp.gate_proj), up_in=lin(mlp.up_proj), down_in=lin(mlp.down_proj))
            if all(v is not None for v in fmts.values()):
                plan = fmts
        self._plan_cache = (seq_len, plan)
        return plan

    def _gated_mlp_operand(self, xg, xu, plan, rows):
        """Q_down(silu(gate_proj(x)) * up_proj(x)) as bf16 [rows, I] (reference modeling_llama.py:84).  gate_proj and up_proj with the
        same x-quantizer read ONE operand: both GEMMs run as one launch over the interleaved weights and the SiLU, the product and
        down_proj's x-quantizer sit in its epilogue (gated_silu_prequantized).  Otherwise: two GEMMs + the silu*mul quantizer kernel."""
        mlp = self.mlp
        if xg.data_ptr() == xu.data_ptr() and gated_silu_fusable(mlp.gate_proj, mlp.up_proj):
            return gated_silu_prequantized(mlp.gate_proj, mlp.up_proj, xg.view(rows, -1), plan["down_in"])
        g = mlp.gate_proj.forward_prequantized(xg)
        u = mlp.up_proj.forward_prequantized(xu)
        return silu_mul_quantize(g.view(rows, -1), u.view(rows, -1), plan["down_in"])

    @torch.no_grad()
    def _split_forward(self, h, position_ids, plan, default_positions=False, key_mask=None):
        """Layer forward for configs whose matmuls keep an fp32 operand (block_log): same fused glue, attention through
        rope_quantize_split / split_attention (scores cross HBM once as fp32 and once as quantised bf16 probabilities)."""
        B, S, H = h.shape
        at, mlp = self.self_attn, self.mlp
        n1, n2 = self.input_layernorm, self.post_attention_layernorm
        qc = at.quant_config
        xq, xk, xv = norm_quantize(h, n1.weight, None, n1.variance_epsilon, [plan["q_in"], plan["k_in"], plan["v_in"]])
        if xq.data_ptr() == xk.data_ptr() == xv.data_ptr() and qkv_plain_fusable(at.q_proj, at.k_proj, at.v_proj):
            q, k, v = qkv_plain_prequantized(at.q_proj, at.k_proj, at.v_proj, xq.view(B * S, H))      # one launch, column views
        else:
            q = at.q_proj.forward_prequantized(xq)
            k = at.k_proj.forward_prequantized(xk)
            v = at.v_proj.forward_prequantized(xv)
        cos, sin = at.rotary_emb(h, seq_len=S)
        Qq, Kp = rope_quantize_split(q.view(B, S, H), k.view(B, S, H), cos, sin, None if default_positions else position_ids,
                                     qc["rotary_positional_encoding"], qc["matmul_0"], at.num_heads)
        o = split_attention(Qq, Kp, v.view(B, S, H), qc["matmul_1"], at.num_heads, math.sqrt(at.head_dim), causal=True, key_mask=key_mask)
        okind, okw = plan["o_in"]
        oq = quantize_operand_bf16(o.view(B * S, H), okind, okw, [1, 16], True)      # o_proj's x-quantizer (exact, incl. the global-min rule)
        h2 = at.o_proj.forward_prequantized(oq, residual=h)
        xg, xu = norm_quantize(h2, n2.weight, None, n2.variance_epsilon, [plan["gate_in"], plan["up_in"]])
        a = self._gated_mlp_operand(xg, xu, plan, B * S)
        h3 = mlp.down_proj.forward_prequantized(a, residual=h2.view(B * S, H))
        return h3.view(B, S, H)

    @torch.no_grad()
    def _fused_forward(self, h, position_ids, plan, default_positions=False, key_mask=None):
        if plan.get("mode") == "split":
            return self._split_forward(h, position_ids, plan, default_positions, key_mask)
        B, S, H = h.shape
        at, mlp = self.self_attn, self.mlp
        n1, n2 = self.input_layernorm, self.post_attention_layernorm
        qc = at.quant_config
        xq, xk, xv = norm_quantize(h, n1.weight, None, n1.variance_epsilon, [plan["q_in"], plan["k_in"], plan["v_in"]])
        cos, sin = at.rotary_emb(h, seq_len=S)
        pos_arg = None if default_positions else position_ids                    # default arange: the kernels derive it (graph-capturable)
        fusedqk, Vq = None, None
        if self.fused_rope and rope_epilogue_fusable(at.q_proj, at.head_dim) and rope_epilogue_fusable(at.k_proj, at.head_dim):
            # RoPE + matmul_0's operand quantizers inside the q_proj / k_proj GEMM epilogues: no fp32 q / k in HBM at all
            prep = rope_quantize_operands(cos, sin, pos_arg, qc["rotary_positional_encoding"], qc["matmul_0"], B, S, at.head_dim)
            if prep is not None:
                cos_t, sin_t, pos, fq, fk = prep
                if (xq.data_ptr() == xk.data_ptr() == xv.data_ptr() and (B * S) % 16 == 0
                        and qkv_rope_fusable(at.q_proj, at.k_proj, at.v_proj, at.head_dim)):
                    # one x-quantizer for the three projections: ONE launch over the concatenated weights
                    Qf, Kf, Vq = qkv_rope_prequantized(at.q_proj, at.k_proj, at.v_proj, xq, cos_t, sin_t, pos, fq, fk, plan["v_out"], S, at.head_dim)
                    fusedqk = (Qf.view(B, S, H), Kf.view(B, S, H))
                else:
                    fusedqk = (rope_prequantized(at.q_proj, xq, cos_t, sin_t, pos, fq, S, at.head_dim, False).view(B, S, H),
                               rope_prequantized(at.k_proj, xk, cos_t, sin_t, pos, fk, S, at.head_dim, True).view(B, S, H))
        if Vq is None:
            Vq = at.v_proj.forward_prequantized(xv, out_format=plan["v_out"])    # bmm_1's y-quantizer in the GEMM epilogue
        if fusedqk is None:
            q = at.q_proj.forward_prequantized(xq)                               # fp32: RoPE runs on unquantised projections
            k = at.k_proj.forward_prequantized(xk)
            if self.fused_rope:                                                  # RoPE + both matmul_0 operand quantizers: 2 kernels
                fusedqk = _rope_token_major_quantized(q.view(B, S, H), k.view(B, S, H), cos, sin, pos_arg,
                                                      qc["rotary_positional_encoding"], qc["matmul_0"], at.num_heads)
        if fusedqk is not None:
            Qq, Kq = fusedqk
        else:
            q4, k4 = _rope_token_major(q.view(B, S, at.num_heads, at.head_dim), k.view(B, S, at.num_heads, at.head_dim), cos, sin,
                                       position_ids, qc["rotary_positional_encoding"])
            Qq, Kq, _ = quantize_qkv(q4.view(B, S, H), k4.view(B, S, H), None, qc["matmul_0"], qc["matmul_1"], at.num_heads)
        oq = fused_causal_attention_q(Qq, Kq, Vq, qc["matmul_1"], at.num_heads, B, S, math.sqrt(at.head_dim),
                                      out_cfg=at.o_proj.config, key_mask=key_mask)
        h2 = at.o_proj.forward_prequantized(oq, residual=h)                      # residual + o_proj(attn)
        xg, xu = norm_quantize(h2, n2.weight, None, n2.variance_epsilon, [plan["gate_in"], plan["up_in"]])
        a = self._gated_mlp_operand(xg, xu, plan, B * S)                         # Q_down(silu(gate) * up)
        h3 = mlp.down_proj.forward_prequantized(a, residual=h2.view(B * S, H))   # residual + down(silu(gate) * up)
        return h3.view(B, S, H)

    def forward(self, hidden_states, attention_mask=None, position_ids=None, output_attentions=False, causal_only=False,
                default_positions=False, key_mask=None):
        if (causal_only and not output_attentions and hidden_states.is_cuda and hidden_states.dtype == torch.float32
                and not torch.is_grad_enabled() and not self.training and hidden_states.ndim == 3):
            plan = self._fused_plan(hidden_states.shape[1])
            if plan is not None:
                return self._fused_forward(hidden_states, position_ids, plan, default_positions, key_mask), None
        residual = hidden_states
        hidden_states = self.input_layernorm(hidden_states)
        hidden_states, attn = self.self_attn(hidden_states, attention_mask=attention_mask, position_ids=position_ids,
                                             output_attentions=output_attentions)
        hidden_states = residual + hidden_states
        residual = hidden_states
        hidden_states = self.post_attention_layernorm(hidden_states)
        hidden_states = self.mlp(hidden_states)
        return residual + hidden_states, attn


class LlamaQuantizedPreTrainedModel(PreTrainedModel):
    config_class = LlamaQuantizedConfig
    config: LlamaQuantizedConfig
    base_model_prefix = "model"
    supports_gradient_checkpointing = False
    _no_split_modules = ["LlamaQuantizedDecoderLayer"]

    @torch.no_grad()
    def _init_weights(self, module):
        std = self.config.initializer_range
        if isinstance(module, nn.Linear):
            module.weight.normal_(mean=0.0, std=std)
            if module.bias is not None:
                module.bias.zero_()
        elif isinstance(module, nn.Embedding):
            module.weight.normal_(mean=0.0, std=std)
            if module.padding_idx is not None:
                module.weight[module.padding_idx].zero_()
        elif isinstance(module, LlamaRMSNorm):
            module.weight.fill_(1.0)


def _causal_mask(attention_mask, bsz, q_len, dtype, device):
    neg = torch.finfo(dtype).min
    mask = torch.triu(torch.full((q_len, q_len), neg, device=device, dtype=dtype), diagonal=1)[None, None].expand(bsz, 1, q_len, q_len)
    if attention_mask is not None and not bool(attention_mask.all()):
        inv = 1.0 - attention_mask[:, None, None, :].to(dtype)
        mask = mask + inv.masked_fill(inv.to(torch.bool), neg)
    return mask


class LlamaQuantizedModel(LlamaQuantizedPreTrainedModel):
    def __init__(self, config: LlamaQuantizedConfig):
        super().__init__(config)
        self.padding_idx = config.pad_token_id
        self.vocab_size = config.vocab_size
        self.embed_tokens = nn.Embedding(config.vocab_size, config.hidden_size, self.padding_idx)
        self.layers = nn.ModuleList([LlamaQuantizedDecoderLayer(config, i) for i in range(config.num_hidden_layers)])
        self.norm = LlamaRMSNorm(config.hidden_size, eps=config.rms_norm_eps)
        self.fused_glue = True           # set False to force the op-by-op module path (QUANTIZED_MODULE_MAP / QUANTIZED_FUNC_MAP)
        self.post_init()

    def get_input_embeddings(self):
        return self.embed_tokens

    def set_input_embeddings(self, value):
        self.embed_tokens = value

    def forward(self, input_ids=None, attention_mask=None, position_ids=None, inputs_embeds=None, output_attentions=False,
                output_hidden_states=False):
        if inputs_embeds is None:
            inputs_embeds = self.embed_tokens(input_ids)
        bsz, q_len = inputs_embeds.shape[:2]
        default_positions = position_ids is None
        if default_positions:
            position_ids = torch.arange(q_len, dtype=torch.long, device=inputs_embeds.device).unsqueeze(0).view(-1, q_len)
        mask = _causal_mask(attention_mask, bsz, q_len, inputs_embeds.dtype, inputs_embeds.device)
        no_padding = attention_mask is None or bool(attention_mask.all())
        # right-padded batches keep the fused layers: the attention kernel takes the key-padding bitmap (every causal row keeps
        # key 0); left padding has fully masked rows -> op-by-op path (see attention.causal_key_mask)
        key_mask = None
        if self.fused_glue and not no_padding and inputs_embeds.is_cuda:
            key_mask = causal_key_mask(attention_mask)
        causal_only = self.fused_glue and (no_padding or key_mask is not None)
        hidden_states = inputs_embeds
        all_h, all_a = (), ()
        for layer in self.layers:
            if output_hidden_states:
                all_h += (hidden_states,)
            hidden_states, attn = layer(hidden_states, attention_mask=mask, position_ids=position_ids,
                                        output_attentions=output_attentions, causal_only=causal_only,
                                        default_positions=default_positions, key_mask=key_mask)
            if output_attentions:
                all_a += (attn,)
        hidden_states = self.norm(hidden_states)
        if output_hidden_states:
            all_h += (hidden_states,)
        return hidden_states, (all_h if output_hidden_states else None), (all_a if output_attentions else None)


class LlamaQuantizedForCausalLM(LlamaQuantizedPreTrainedModel):
    _tied_weights_keys = {"lm_head.weight": "model.embed_tokens.weight"}

    def __init__(self, config: LlamaQuantizedConfig):
        super().__init__(config)
        self.model = LlamaQuantizedModel(config)
        self.lm_head = nn.Linear(config.hidden_size, config.vocab_size, bias=False)   # unquantised (reference :772)
        self.post_init()

    def get_input_embeddings(self):
        return self.model.embed_tokens

    def set_input_embeddings(self, value):
        self.model.embed_tokens = value

    def get_output_embeddings(self):
        return self.lm_head

    def set_output_embeddings(self, new_embeddings):
        self.lm_head = new_embeddings

    def forward(self, input_ids=None, attention_mask=None, position_ids=None, inputs_embeds=None, labels=None,
                output_attentions=False, output_hidden_states=False, return_dict=True, **unused):
        hidden, all_h, all_a = self.model(input_ids=input_ids, attention_mask=attention_mask, position_ids=position_ids,
                                          inputs_embeds=inputs_embeds, output_attentions=output_attentions,
                                          output_hidden_states=output_hidden_states)
        # unquantised fp32 head (reference keeps nn.Linear): fp32-equivalent split-bf16 GEMM on the tensor cores
        logits = fp32_linear(hidden, self.lm_head.weight, self.lm_head.bias)
        loss = None
        if labels is not None:
            # shifted CE (modeling_llama.py:867-879) in one streaming read of the logits: bq_token_ce_mean
            loss = causal_lm_loss(logits, labels, shift=True)
        if not return_dict:
            out = (logits, None, all_h, all_a)
            return ((loss,) + out) if loss is not None else out
        return CausalLMOutputWithPast(loss=loss, logits=logits, past_key_values=None, hidden_states=all_h, attentions=all_a)


class LlamaQuantizedForSequenceClassification(LlamaQuantizedPreTrainedModel):
    """reference modeling_llama.py:955-1080 — MODEL_MAP["llama"]["cls"]; `score` is an unquantised bias-free nn.Linear."""

    def __init__(self, config: LlamaQuantizedConfig):
        super().__init__(config)
        self.num_labels = config.num_labels
        self.model = LlamaQuantizedModel(config)
        self.score = nn.Linear(config.hidden_size, self.num_labels, bias=False)
        self.post_init()

    def get_input_embeddings(self):
        return self.model.embed_tokens

    def set_input_embeddings(self, value):
        self.model.embed_tokens = value

    def forward(self, input_ids=None, attention_mask=None, position_ids=None, inputs_embeds=None, labels=None,
                output_attentions=False, output_hidden_states=False, return_dict=True, **unused):
        from ..opt_quantized.modeling_opt import sequence_classification_head

        hidden, all_h, all_a = self.model(input_ids=input_ids, attention_mask=attention_mask, position_ids=position_ids,
                                          inputs_embeds=inputs_embeds, output_attentions=output_attentions,
                                          output_hidden_states=output_hidden_states)
        pooled, loss = sequence_classification_head(self.config, self.score(hidden), input_ids, inputs_embeds, labels, self.num_labels)
        if not return_dict:
            out = (pooled, None, all_h, all_a)
            return ((loss,) + out) if loss is not None else out
        return SequenceClassifierOutputWithPast(loss=loss, logits=pooled, past_key_values=None, hidden_states=all_h, attentions=all_a)
