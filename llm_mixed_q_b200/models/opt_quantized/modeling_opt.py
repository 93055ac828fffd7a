"""
Quantized OPT — module classes with HF-compatible parameter names
(`model.decoder.layers.<i>.self_attn.q_proj.weight`, …) so OPT checkpoints load unchanged.

Mirrors the quantisation wiring of reference opt_quantized/modeling_opt.py: six quantized Linears and two
quantized bmm per decoder layer (:174-177, :246, :312, :353-354), everything else (embeddings, LayerNorm,
softmax, ReLU, residuals, lm_head, loss) unquantised fp32 (:831-832, :942-944, :1086-1098).  The
forward-only tier supports the perplexity path (`input_ids`, `attention_mask`, `labels`); KV-cache decoding
is not wired.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
from torch.nn import CrossEntropyLoss
from transformers.activations import ACT2FN
from transformers.modeling_outputs import CausalLMOutputWithPast, SequenceClassifierOutputWithPast
from transformers.modeling_utils import PreTrainedModel

from ..quantize import get_quantized_cls, get_quantized_func
from ..quantize.quantized_functions.fp32_linear import fp32_linear
from ..quantize.quantized_functions.loss import causal_lm_loss
from ..quantize.quantized_functions.attention import (causal_key_mask, fusable as _attn_fusable, fused_causal_attention,
                                                      fused_causal_attention_q, output_quantizable)
from ..quantize.quantized_functions.fused_glue import _NORM_KINDS, linear_input_format, norm_quantize, row_block16_format
from ..quantize.quantized_functions.split_attention import rope_quantize_split, split_attention, splittable as _attn_splittable
from ..quantize.quantized_modules.linear import quantize_operand_bf16
from .configuration_opt import OPTQuantizedConfig


class OPTLearnedPositionalEmbedding(nn.Embedding):
    """positions offset by 2 (HF OPT convention; reference modeling_opt.py:112-140)."""

    def __init__(self, num_embeddings: int, embedding_dim: int):
        self.offset = 2
        super().__init__(num_embeddings + self.offset, embedding_dim)

    def forward(self, attention_mask: torch.LongTensor, past_key_values_length: int = 0):
        attention_mask = attention_mask.long()
        positions = (torch.cumsum(attention_mask, dim=1).type_as(attention_mask) * attention_mask).long() - 1
        positions = positions[:, past_key_values_length:]
        return super().forward(positions + self.offset)


def _causal_additive_mask(attention_mask, bsz, tgt_len, dtype, device, all_ones=None):
    """[bsz, 1, tgt, src] additive mask: finfo.min above the diagonal and on padded keys (reference :520-548).
    `all_ones`: the caller already knows whether the mask has no padding (None: look — a device-to-host synchronisation)."""
    neg = torch.finfo(dtype).min
    mask = torch.full((tgt_len, tgt_len), neg, device=device, dtype=dtype)
    mask = torch.triu(mask, diagonal=1)[None, None].expand(bsz, 1, tgt_len, tgt_len)
    if all_ones is None:
        all_ones = attention_mask is None or bool(attention_mask.all())
    if attention_mask is not None and not all_ones:
        pad = (1.0 - attention_mask[:, None, None, :].to(dtype)).masked_fill(attention_mask[:, None, None, :] == 0, 1.0)
        pad = pad.masked_fill(pad.bool(), neg)
        mask = mask + pad
        mask = mask.clamp(min=neg)
    return mask


class OPTQauntizedAttention(nn.Module):      # (sic) class name kept from the reference, modeling_opt.py:143
    def __init__(self, embed_dim: int, num_heads: int, dropout: float = 0.0, is_decoder: bool = False, bias: bool = True,
                 quant_config: dict = None):
        super().__init__()
        self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.dropout = dropout
        self.head_dim = embed_dim // num_heads
        if self.head_dim * num_heads != embed_dim:
            raise ValueError(f"embed_dim must be divisible by num_heads (got `embed_dim`: {embed_dim} and `num_heads`: {num_heads}).")
        self.scaling = self.head_dim**-0.5
        self.is_decoder = is_decoder
        qc = quant_config
        self.k_proj = get_quantized_cls("linear", qc["k_proj"])(embed_dim, embed_dim, bias=bias, config=qc["k_proj"])
        self.q_proj = get_quantized_cls("linear", qc["q_proj"])(embed_dim, embed_dim, bias=bias, config=qc["q_proj"])
        self.v_proj = get_quantized_cls("linear", qc["v_proj"])(embed_dim, embed_dim, bias=bias, config=qc["v_proj"])
        self.out_proj = get_quantized_cls("linear", qc["out_proj"])(embed_dim, embed_dim, bias=bias, config=qc["out_proj"])
        self.quant_config = qc

    def _shape(self, tensor: torch.Tensor, seq_len: int, bsz: int):
        return tensor.view(bsz, seq_len, self.num_heads, self.head_dim).transpose(1, 2).contiguous()

    def forward(self, hidden_states: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                output_attentions: bool = False, causal_only: bool = False, key_mask: Optional[torch.Tensor] = None):
        """`causal_only`: the additive mask is the causal one, plus — when `key_mask` (attention.key_mask_bits) is given — key
        padding that leaves every query row at least one key; the fused kernel then applies both itself."""
        bsz, tgt_len, _ = hidden_states.size()
        if (causal_only and not output_attentions and hidden_states.is_cuda and not torch.is_grad_enabled()
                and not (self.training and self.dropout > 0)
                and _attn_fusable(self.quant_config["bmm_0"], self.quant_config["bmm_1"], self.head_dim, tgt_len)):
            # one kernel for bmm_0 -> causal mask -> softmax -> bmm_1; scores / probabilities stay on chip
            q = self.q_proj(hidden_states) * self.scaling
            k = self.k_proj(hidden_states)
            v = self.v_proj(hidden_states)
            if self.out_proj.accepts_prequantized() and output_quantizable(self.out_proj.config, self.embed_dim, tgt_len):
                # the x-quantizer of out_proj runs in the attention epilogue; out_proj consumes the bf16 operand directly
                oq = fused_causal_attention(q, k, v, self.quant_config["bmm_0"], self.quant_config["bmm_1"], self.num_heads,
                                            score_div=1.0, out_cfg=self.out_proj.config, key_mask=key_mask)
                return self.out_proj.forward_prequantized(oq), None
            attn_output = fused_causal_attention(q, k, v, self.quant_config["bmm_0"], self.quant_config["bmm_1"],
                                                 self.num_heads, score_div=1.0, key_mask=key_mask)
            return self.out_proj(attn_output), None
        query_states = self.q_proj(hidden_states) * self.scaling
        key_states = self._shape(self.k_proj(hidden_states), -1, bsz)
        value_states = self._shape(self.v_proj(hidden_states), -1, bsz)

        proj_shape = (bsz * self.num_heads, -1, self.head_dim)
        query_states = self._shape(query_states, tgt_len, bsz).view(*proj_shape)
        key_states = key_states.view(*proj_shape)
        value_states = value_states.view(*proj_shape)
        src_len = key_states.size(1)

        bmm_0 = get_quantized_func("bmm", self.quant_config["bmm_0"])
        attn_weights = bmm_0(query_states, key_states.transpose(1, 2), config=self.quant_config["bmm_0"])

        if attention_mask is not None:
            attn_weights = attn_weights.view(bsz, self.num_heads, tgt_len, src_len) + attention_mask
            attn_weights = torch.max(attn_weights, torch.tensor(torch.finfo(attn_weights.dtype).min, device=attn_weights.device))
            attn_weights = attn_weights.view(bsz * self.num_heads, tgt_len, src_len)
        attn_weights = nn.functional.softmax(attn_weights, dim=-1)
        attn_probs = nn.functional.dropout(attn_weights, p=self.dropout, training=self.training)

        bmm_1 = get_quantized_func("bmm", self.quant_config["bmm_1"])
        attn_output = bmm_1(attn_probs, value_states, config=self.quant_config["bmm_1"])

        attn_output = attn_output.view(bsz, self.num_heads, tgt_len, self.head_dim).transpose(1, 2)
        attn_output = attn_output.reshape(bsz, tgt_len, self.embed_dim)
        attn_output = self.out_proj(attn_output)
        return attn_output, (attn_weights.view(bsz, self.num_heads, tgt_len, src_len) if output_attentions else None)


class OPTQuantizedDecoderLayer(nn.Module):
    def __init__(self, config: OPTQuantizedConfig, layer_id: int):
        super().__init__()
        self.embed_dim = config.hidden_size
        qc = config.quant_config[f"model_layer_{layer_id}"]
        self.self_attn = OPTQauntizedAttention(embed_dim=self.embed_dim, num_heads=config.num_attention_heads,
                                               dropout=config.attention_dropout, is_decoder=True, bias=config.enable_bias,
                                               quant_config=qc["self_attn"])
        self.do_layer_norm_before = config.do_layer_norm_before
        self.dropout = config.dropout
        self.activation_fn = ACT2FN[config.activation_function]
        self.activation_name = config.activation_function
        self.self_attn_layer_norm = nn.LayerNorm(self.embed_dim, elementwise_affine=config.layer_norm_elementwise_affine)
        self.fc1 = get_quantized_cls("linear", qc["fc1"])(self.embed_dim, config.ffn_dim, bias=config.enable_bias, config=qc["fc1"])
        self.fc2 = get_quantized_cls("linear", qc["fc2"])(config.ffn_dim, self.embed_dim, bias=config.enable_bias, config=qc["fc2"])
        self.final_layer_norm = nn.LayerNorm(self.embed_dim, elementwise_affine=config.layer_norm_elementwise_affine)

    # ------------------------------------------------------------------ fused layer (PTQ inference, causal mask)
    def _fused_plan(self, seq_len: int):
        """Formats for running the whole layer on 8 fused kernels, or None when any piece is not eligible.
        Each x-quantizer then runs inside the kernel that produces its input: LN->quantize, GEMM epilogues
        (bias, q scaling, ReLU, residual add, quantize), attention epilogue; every GEMM reads a bf16 operand."""
        key = seq_len
        cache = getattr(self, "_plan_cache", None)
        if cache is not None and cache[0] == key:
            return cache[1]
        at = self.self_attn
        plan = None
        H, d = self.embed_dim, at.head_dim
        common = (self.do_layer_norm_before and getattr(self, "activation_name", None) == "relu" and seq_len % 16 == 0
                  and H % 32 == 0 and self.fc1.out_features % 32 == 0
                  and self.self_attn_layer_norm.elementwise_affine and self.final_layer_norm.elementwise_affine)
        ok = common and _attn_fusable(at.quant_config["bmm_0"], at.quant_config["bmm_1"], d, seq_len)
        if common and not ok and _attn_splittable(at.quant_config["bmm_0"], at.quant_config["bmm_1"], d, seq_len):
            # block_log: bmm_0 / bmm_1 keep k / v in fp32 (reference matmul.py:286-297) -> three-kernel attention
            # (split_attention.py); the Linears' x-quantizers run inside LayerNorm / the fc1 epilogue (carrier rule, bq.h)
            lin = lambda m, rows: linear_input_format(m, rows=rows, kinds=_NORM_KINDS)
            fmts = dict(mode="split", q_in=lin(at.q_proj, seq_len), k_in=lin(at.k_proj, seq_len), v_in=lin(at.v_proj, seq_len),
                        o_in=lin(at.out_proj, seq_len), fc1_in=lin(self.fc1, 1), fc2_in=lin(self.fc2, 1))
            if all(v is not None for v in fmts.values()):
                plan = fmts
        if ok:
            fmts = dict(
                # q/k/v/out_proj see the 3-D [B, S, H] activation (reference modeling_opt.py:206-225, :327), fc1 / fc2 the
                # 2-D [B*S, H] one (:412); the bmm operands are [B*h, S, d] / [B*h, d, S]
                q_in=linear_input_format(at.q_proj, rows=seq_len), k_in=linear_input_format(at.k_proj, rows=seq_len),
                v_in=linear_input_format(at.v_proj, rows=seq_len),
                o_in=linear_input_format(at.out_proj, rows=seq_len) if output_quantizable(at.out_proj.config, H, seq_len) else None,
                fc1_in=linear_input_format(self.fc1), fc2_in=linear_input_format(self.fc2),
                q_out=row_block16_format(at.quant_config["bmm_0"], "data_in", d, rows=seq_len),
                k_out=row_block16_format(at.quant_config["bmm_0"], "weight", seq_len, rows=d),
                v_out=row_block16_format(at.quant_config["bmm_1"], "weight", d, rows=seq_len))
            if all(v is not None for v in fmts.values()):
                plan = fmts
        self._plan_cache = (key, plan)
        return plan

    @torch.no_grad()
    def _split_forward(self, h, plan, key_mask=None):
        """Layer forward for configs whose bmms keep an fp32 operand (block_log): fused glue around the three-kernel attention."""
        B, S, H = h.shape
        at = self.self_attn
        ln1, ln2 = self.self_attn_layer_norm, self.final_layer_norm
        xq_q, xq_k, xq_v = norm_quantize(h, ln1.weight, ln1.bias, ln1.eps, [plan["q_in"], plan["k_in"], plan["v_in"]])
        q = at.q_proj.forward_prequantized(xq_q, scale=at.scaling)                    # q_proj(x) * scaling, then bmm_0's x-quantizer
        k = at.k_proj.forward_prequantized(xq_k)
        v = at.v_proj.forward_prequantized(xq_v)
        Qq, Kp = rope_quantize_split(q.view(B, S, H), k.view(B, S, H), None, None, None, None, at.quant_config["bmm_0"], at.num_heads)
        o = split_attention(Qq, Kp, v.view(B, S, H), at.quant_config["bmm_1"], at.num_heads, 1.0, causal=True, key_mask=key_mask)
        okind, okw = plan["o_in"]
        oq = quantize_operand_bf16(o.view(B * S, H), okind, okw, [1, 16], True)       # out_proj's x-quantizer (exact)
        h2 = at.out_proj.forward_prequantized(oq, residual=h)
        (x1,) = norm_quantize(h2, ln2.weight, ln2.bias, ln2.eps, [plan["fc1_in"]])
        a = self.fc1.forward_prequantized(x1.view(B * S, H), relu=True, out_format=plan["fc2_in"])
        h3 = self.fc2.forward_prequantized(a, residual=h2.view(B * S, H))
        return h3.view(B, S, H)

    @torch.no_grad()
    def _fused_forward(self, h, plan, key_mask=None):
        if plan.get("mode") == "split":
            return self._split_forward(h, plan, key_mask)
        B, S, H = h.shape
        at = self.self_attn
        ln1, ln2 = self.self_attn_layer_norm, self.final_layer_norm
        xq_q, xq_k, xq_v = norm_quantize(h, ln1.weight, ln1.bias, ln1.eps, [plan["q_in"], plan["k_in"], plan["v_in"]])
        Qq = at.q_proj.forward_prequantized(xq_q, scale=at.scaling, out_format=plan["q_out"])
        Kq = at.k_proj.forward_prequantized(xq_k, out_format=plan["k_out"], out_blocks_along_rows=True)
        Vq = at.v_proj.forward_prequantized(xq_v, out_format=plan["v_out"])
        oq = fused_causal_attention_q(Qq, Kq, Vq, at.quant_config["bmm_1"], at.num_heads, B, S, 1.0, out_cfg=at.out_proj.config,
                                      key_mask=key_mask)
        h2 = at.out_proj.forward_prequantized(oq, residual=h)                         # residual + out_proj(attn)
        (x1,) = norm_quantize(h2, ln2.weight, ln2.bias, ln2.eps, [plan["fc1_in"]])
        a = self.fc1.forward_prequantized(x1.view(B * S, H), relu=True, out_format=plan["fc2_in"])
        h3 = self.fc2.forward_prequantized(a, residual=h2.view(B * S, H))             # residual + fc2(relu(fc1(x)))
        return h3.view(B, S, H)

    def forward(self, hidden_states, attention_mask=None, output_attentions=False, causal_only=False, fused_glue=False,
                key_mask=None):
        if (fused_glue and causal_only and not output_attentions and hidden_states.is_cuda and hidden_states.dtype == torch.float32
                and not torch.is_grad_enabled() and not self.training and hidden_states.ndim == 3):
            plan = self._fused_plan(hidden_states.shape[1])
            if plan is not None:
                return self._fused_forward(hidden_states, plan, key_mask), None
        residual = hidden_states
        if self.do_layer_norm_before:
            hidden_states = self.self_attn_layer_norm(hidden_states)
        hidden_states, attn = self.self_attn(hidden_states, attention_mask=attention_mask, output_attentions=output_attentions,
                                             causal_only=causal_only, key_mask=key_mask)
        hidden_states = nn.functional.dropout(hidden_states, p=self.dropout, training=self.training)
        hidden_states = residual + hidden_states
        if not self.do_layer_norm_before:
            hidden_states = self.self_attn_layer_norm(hidden_states)

        shape = hidden_states.shape
        hidden_states = hidden_states.reshape(-1, hidden_states.size(-1))      # fc1/fc2 see a 2-D input (reference :412)
        residual = hidden_states
        if self.do_layer_norm_before:
            hidden_states = self.final_layer_norm(hidden_states)
        hidden_states = self.fc2(self.activation_fn(self.fc1(hidden_states)))
        hidden_states = nn.functional.dropout(hidden_states, p=self.dropout, training=self.training)
        hidden_states = (residual + hidden_states).view(shape)
        if not self.do_layer_norm_before:
            hidden_states = self.final_layer_norm(hidden_states)
        return hidden_states, attn


class OPTQuantizedPreTrainedModel(PreTrainedModel):
    config_class = OPTQuantizedConfig
    config: OPTQuantizedConfig
    base_model_prefix = "model"
    supports_gradient_checkpointing = False
    _no_split_modules = ["OPTQuantizedDecoderLayer"]

    @torch.no_grad()
    def _init_weights(self, module):
        std = self.config.init_std                  # reference modeling_opt.py:472-481
        if isinstance(module, nn.Linear):
            module.weight.normal_(mean=0.0, std=std)
            if module.bias is not None:
                module.bias.zero_()
        elif isinstance(module, nn.Embedding):
            module.weight.normal_(mean=0.0, std=std)
            if module.padding_idx is not None:
                module.weight[module.padding_idx].zero_()
        elif isinstance(module, nn.LayerNorm):
            if module.weight is not None:
                module.weight.fill_(1.0)
                module.bias.zero_()


class OPTQuantizedDecoder(OPTQuantizedPreTrainedModel):
    def __init__(self, config: OPTQuantizedConfig):
        super().__init__(config)
        self.dropout = config.dropout
        self.padding_idx = config.pad_token_id
        self.max_target_positions = config.max_position_embeddings
        self.vocab_size = config.vocab_size
        self.embed_tokens = nn.Embedding(config.vocab_size, config.word_embed_proj_dim, self.padding_idx)
        self.embed_positions = OPTLearnedPositionalEmbedding(config.max_position_embeddings, config.hidden_size)
        if config.word_embed_proj_dim != config.hidden_size:
            self.project_out = nn.Linear(config.hidden_size, config.word_embed_proj_dim, bias=False)
            self.project_in = nn.Linear(config.word_embed_proj_dim, config.hidden_size, bias=False)
        else:
            self.project_out = None
            self.project_in = None
        if config.do_layer_norm_before and not config._remove_final_layer_norm:
            self.final_layer_norm = nn.LayerNorm(config.hidden_size, elementwise_affine=config.layer_norm_elementwise_affine)
        else:
            self.final_layer_norm = None
        self.layers = nn.ModuleList([OPTQuantizedDecoderLayer(config, i) for i in range(config.num_hidden_layers)])
        self.fused_attention = True      # set False to force the op-by-op attention path (QUANTIZED_FUNC_MAP bmm functions)
        self.fused_glue = True           # set False to run LayerNorm / ReLU / residual / x-quantizers as separate kernels
        self.post_init()

    def get_input_embeddings(self):
        return self.embed_tokens

    def set_input_embeddings(self, value):
        self.embed_tokens = value

    def forward(self, input_ids=None, attention_mask=None, inputs_embeds=None, output_attentions=False,
                output_hidden_states=False):
        if inputs_embeds is None:
            inputs_embeds = self.embed_tokens(input_ids)
        bsz, seq_len = inputs_embeds.shape[:2]
        # no padding: known without looking when no mask was passed (the look is a device-to-host synchronisation that would drain
        # the launch queue at the start of every forward and cannot be captured in a CUDA graph)
        no_padding = attention_mask is None or bool(attention_mask.all())
        if attention_mask is None:
            attention_mask = torch.ones(bsz, seq_len, dtype=torch.bool, device=inputs_embeds.device)
        # padded batches stay on the fused kernels when every causal row keeps a key (key 0 valid: right padding) — the kernel
        # takes the key-padding bitmap; left padding has fully masked rows, which the reference turns into a uniform
        # distribution over ALL keys (:520-548 + the max(finfo.min) clamp): op-by-op path
        key_mask = None
        if self.fused_attention and not no_padding and inputs_embeds.is_cuda:
            key_mask = causal_key_mask(attention_mask)
        causal_only = self.fused_attention and (no_padding or key_mask is not None)
        causal = _causal_additive_mask(attention_mask, bsz, seq_len, inputs_embeds.dtype, inputs_embeds.device, all_ones=no_padding)
        pos_embeds = self.embed_positions(attention_mask, 0)
        if self.project_in is not None:
            inputs_embeds = self.project_in(inputs_embeds)
        hidden_states = inputs_embeds + pos_embeds
        hidden_states = nn.functional.dropout(hidden_states, p=self.dropout, training=self.training)
        all_h, all_a = ((), ())
        for layer in self.layers:
            if output_hidden_states:
                all_h += (hidden_states,)
            hidden_states, attn = layer(hidden_states, attention_mask=causal, output_attentions=output_attentions,
                                        causal_only=causal_only, fused_glue=self.fused_glue and causal_only, key_mask=key_mask)
            if output_attentions:
                all_a += (attn,)
        if self.final_layer_norm is not None:
            hidden_states = self.final_layer_norm(hidden_states)
        if self.project_out is not None:
            hidden_states = self.project_out(hidden_states)
        if output_hidden_states:
            all_h += (hidden_states,)
        return hidden_states, (all_h if output_hidden_states else None), (all_a if output_attentions else None)


class OPTQuantizedModel(OPTQuantizedPreTrainedModel):
    def __init__(self, config: OPTQuantizedConfig):
        super().__init__(config)
        self.decoder = OPTQuantizedDecoder(config)
        self.post_init()

    def get_input_embeddings(self):
        return self.decoder.embed_tokens

    def set_input_embeddings(self, value):
        self.decoder.embed_tokens = value

    def forward(self, *args, **kwargs):
        return self.decoder(*args, **kwargs)


class OPTQuantizedForCausalLM(OPTQuantizedPreTrainedModel):
    _tied_weights_keys = {"lm_head.weight": "model.decoder.embed_tokens.weight"}

    def __init__(self, config: OPTQuantizedConfig):
        super().__init__(config)
        self.model = OPTQuantizedModel(config)
        # unquantised fp32 head, tied to the token embedding (reference modeling_opt.py:942-944)
        self.lm_head = nn.Linear(config.word_embed_proj_dim, config.vocab_size, bias=False)
        self.post_init()

    def get_input_embeddings(self):
        return self.model.decoder.embed_tokens

    def set_input_embeddings(self, value):
        self.model.decoder.embed_tokens = value

    def get_output_embeddings(self):
        return self.lm_head

    def set_output_embeddings(self, new_embeddings):
        self.lm_head = new_embeddings

    def forward(self, input_ids=None, attention_mask=None, inputs_embeds=None, labels=None, output_attentions=False,
                output_hidden_states=False, return_dict=True, **unused):
        hidden, all_h, all_a = self.model.decoder(input_ids=input_ids, attention_mask=attention_mask,
                                                  inputs_embeds=inputs_embeds, output_attentions=output_attentions,
                                                  output_hidden_states=output_hidden_states)
        # unquantised fp32 head (reference keeps nn.Linear): fp32-equivalent split-bf16 GEMM on the tensor cores
        logits = fp32_linear(hidden, self.lm_head.weight, self.lm_head.bias)
        loss = None
        if labels is not None:
            # shifted CE (modeling_opt.py:1086-1098) in one streaming read of the logits: bq_token_ce_mean
            loss = causal_lm_loss(logits, labels, shift=True)
        if not return_dict:
            out = (logits, None, all_h, all_a)
            return ((loss,) + out) if loss is not None else out
        return CausalLMOutputWithPast(loss=loss, logits=logits, past_key_values=None, hidden_states=all_h, attentions=all_a)


def sequence_classification_head(config, logits, input_ids, inputs_embeds, labels, num_labels):
    """Shared by OPT / Llama ForSequenceClassification (reference modeling_opt.py:1221-1273, modeling_llama.py:1010-1063):
    pool the logits of the last non-pad token of every sequence, then HF problem-type dispatch for the loss."""
    from torch.nn import BCEWithLogitsLoss, MSELoss

    batch_size = input_ids.shape[0] if input_ids is not None else inputs_embeds.shape[0]
    if config.pad_token_id is None or input_ids is None:
        sequence_lengths = -1
    else:
        sequence_lengths = (torch.ne(input_ids, config.pad_token_id).sum(-1) - 1).to(logits.device)
    pooled = logits[torch.arange(batch_size, device=logits.device), sequence_lengths]
    loss = None
    if labels is not None:
        if config.problem_type is None:
            if num_labels == 1:
                config.problem_type = "regression"
            elif num_labels > 1 and labels.dtype in (torch.long, torch.int):
                config.problem_type = "single_label_classification"
            else:
                config.problem_type = "multi_label_classification"
        if config.problem_type == "regression":
            loss = MSELoss()(pooled.squeeze(), labels.squeeze()) if num_labels == 1 else MSELoss()(pooled, labels)
        elif config.problem_type == "single_label_classification":
            loss = CrossEntropyLoss()(pooled.view(-1, num_labels), labels.view(-1))
        else:
            loss = BCEWithLogitsLoss()(pooled, labels)
    return pooled, loss


class OPTQuantizedForSequenceClassification(OPTQuantizedPreTrainedModel):
    """reference modeling_opt.py:1163-1290 — MODEL_MAP["opt"]["cls"]; `score` is an unquantised bias-free nn.Linear."""

    def __init__(self, config: OPTQuantizedConfig):
        super().__init__(config)
        self.num_labels = config.num_labels
        self.model = OPTQuantizedModel(config)
        self.score = nn.Linear(config.word_embed_proj_dim, self.num_labels, bias=False)
        self.post_init()

    def get_input_embeddings(self):
        return self.model.decoder.embed_tokens

    def set_input_embeddings(self, value):
        self.model.decoder.embed_tokens = value

    def forward(self, input_ids=None, attention_mask=None, inputs_embeds=None, labels=None, output_attentions=False,
                output_hidden_states=False, return_dict=True, **unused):
        hidden, all_h, all_a = self.model.decoder(input_ids=input_ids, attention_mask=attention_mask,
                                                  inputs_embeds=inputs_embeds, output_attentions=output_attentions,
                                                  output_hidden_states=output_hidden_states)
        pooled, loss = sequence_classification_head(self.config, self.score(hidden), input_ids, inputs_embeds, labels, self.num_labels)
        if not return_dict:
            out = (pooled, None, all_h, all_a)
            return ((loss,) + out) if loss is not None else out
        return SequenceClassifierOutputWithPast(loss=loss, logits=pooled, past_key_values=None, hidden_states=all_h, attentions=all_a)
