"""
Quantized BERT — HF-compatible parameter names (`bert.encoder.layer.<i>.attention.self.query.weight`, …).

Quantisation wiring of reference bert_quantized/modeling_bert.py: six quantized Linears per layer
(query/key/value :281-289, attention.output.dense :454, intermediate.dense :536, output.dense :557), two quantized
4-D matmuls (:366-370 on the k^T view, :433-435; scores divided by sqrt(d) AFTER matmul_0, additive mask after that),
fp32 embeddings / LayerNorm (post-LN) / softmax / GELU / pooler / task heads.  Encoder-only, absolute positions
(the reference's own config comment says cross attention is unsupported, quant_config_bert.py:27); forward-only.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn
from torch.nn import BCEWithLogitsLoss, CrossEntropyLoss, MSELoss
from transformers.activations import ACT2FN
from transformers.modeling_outputs import (BaseModelOutputWithPoolingAndCrossAttentions, MaskedLMOutput,
                                           QuestionAnsweringModelOutput, SequenceClassifierOutput, TokenClassifierOutput)
from transformers.modeling_utils import PreTrainedModel

from ..quantize import get_quantized_cls, get_quantized_func
from ..quantize.quantized_functions.attention import fusable as _attn_fusable, fused_causal_attention, key_mask_bits, output_quantizable
from .configuration_bert import BertQuantizedConfig


class BertEmbeddings(nn.Module):
    """word + token-type + absolute position embeddings, LayerNorm, dropout (reference :185-264)."""

    def __init__(self, config):
        super().__init__()
        self.word_embeddings = nn.Embedding(config.vocab_size, config.hidden_size, padding_idx=config.pad_token_id)
        self.position_embeddings = nn.Embedding(config.max_position_embeddings, config.hidden_size)
        self.token_type_embeddings = nn.Embedding(config.type_vocab_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)
        self.position_embedding_type = getattr(config, "position_embedding_type", "absolute")
        self.register_buffer("position_ids", torch.arange(config.max_position_embeddings).expand((1, -1)), persistent=False)

    def forward(self, input_ids=None, token_type_ids=None, position_ids=None, inputs_embeds=None):
        shape = input_ids.size() if input_ids is not None else inputs_embeds.size()[:-1]
        seq = shape[1]
        if position_ids is None:
            position_ids = self.position_ids[:, :seq]
        if token_type_ids is None:
            token_type_ids = torch.zeros(shape, dtype=torch.long, device=position_ids.device)
        if inputs_embeds is None:
            inputs_embeds = self.word_embeddings(input_ids)
        emb = inputs_embeds + self.token_type_embeddings(token_type_ids)
        if self.position_embedding_type == "absolute":
            emb = emb + self.position_embeddings(position_ids)
        return self.dropout(self.LayerNorm(emb))


class BertQuantizedSelfAttention(nn.Module):
    def __init__(self, config, position_embedding_type=None, quant_config: dict = None):
        super().__init__()
        if config.hidden_size % config.num_attention_heads != 0 and not hasattr(config, "embedding_size"):
            raise ValueError(f"The hidden size ({config.hidden_size}) is not a multiple of the number of attention "
                             f"heads ({config.num_attention_heads})")
        self.num_attention_heads = config.num_attention_heads
        self.attention_head_size = int(config.hidden_size / config.num_attention_heads)
        self.all_head_size = self.num_attention_heads * self.attention_head_size
        qc = quant_config
        self.query = get_quantized_cls("linear", qc["query"])(config.hidden_size, self.all_head_size, config=qc["query"])
        self.key = get_quantized_cls("linear", qc["key"])(config.hidden_size, self.all_head_size, config=qc["key"])
        self.value = get_quantized_cls("linear", qc["value"])(config.hidden_size, self.all_head_size, config=qc["value"])
        self.quant_config = qc
        self.dropout = nn.Dropout(config.attention_probs_dropout_prob)
        self.position_embedding_type = position_embedding_type or getattr(config, "position_embedding_type", "absolute")
        if self.position_embedding_type != "absolute":
            raise NotImplementedError("relative position embeddings are outside the quantized hot path")

    def transpose_for_scores(self, x: torch.Tensor) -> torch.Tensor:
        return x.view(x.size()[:-1] + (self.num_attention_heads, self.attention_head_size)).permute(0, 2, 1, 3)

    def fused_eligible(self, hidden_states, key_mask, head_mask, output_attentions) -> bool:
        """matmul_0 -> /sqrt(d) -> +mask -> softmax -> matmul_1 (:366-435) as ONE kernel (bq_attention_masked, bidirectional with
        the key-padding bitmap): scores and probabilities never reach HBM."""
        return (key_mask is not None and head_mask is None and not output_attentions and hidden_states.is_cuda
                and hidden_states.dtype == torch.float32 and hidden_states.ndim == 3 and not torch.is_grad_enabled()
                and not (self.training and self.dropout.p > 0)
                and _attn_fusable(self.quant_config["matmul_0"], self.quant_config["matmul_1"], self.attention_head_size,
                                  hidden_states.shape[1]))

    def fused_forward(self, hidden_states, key_mask, out_cfg=None):
        """fp32 context [B, S, H], or with `out_cfg` (config of attention.output.dense) its bf16 x-quantised form."""
        return fused_causal_attention(self.query(hidden_states), self.key(hidden_states), self.value(hidden_states),
                                      self.quant_config["matmul_0"], self.quant_config["matmul_1"], self.num_attention_heads,
                                      score_div=math.sqrt(self.attention_head_size), out_cfg=out_cfg, causal=False, key_mask=key_mask)

    def forward(self, hidden_states, attention_mask=None, head_mask=None, output_attentions=False, key_mask=None):
        if self.fused_eligible(hidden_states, key_mask, head_mask, output_attentions):
            return (self.fused_forward(hidden_states, key_mask),)
        query_layer = self.transpose_for_scores(self.query(hidden_states))
        key_layer = self.transpose_for_scores(self.key(hidden_states))
        value_layer = self.transpose_for_scores(self.value(hidden_states))
        mm0, mm1 = self.quant_config["matmul_0"], self.quant_config["matmul_1"]
        scores = get_quantized_func("matmul", mm0)(query_layer, key_layer.transpose(-1, -2), config=mm0)
        scores = scores / math.sqrt(self.attention_head_size)
        if attention_mask is not None:
            scores = scores + attention_mask
        probs = self.dropout(nn.functional.softmax(scores, dim=-1))
        if head_mask is not None:
            probs = probs * head_mask
        context = get_quantized_func("matmul", mm1)(probs, value_layer, config=mm1)
        context = context.permute(0, 2, 1, 3).contiguous()
        context = context.view(context.size()[:-2] + (self.all_head_size,))
        return (context, probs) if output_attentions else (context,)


class BertQuantizedSelfOutput(nn.Module):
    def __init__(self, config, quant_config: dict):
        super().__init__()
        qc = quant_config["dense"]
        self.dense = get_quantized_cls("linear", qc)(config.hidden_size, config.hidden_size, config=qc)
        self.quant_config = quant_config
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)

    def forward(self, hidden_states, input_tensor):
        return self.LayerNorm(self.dropout(self.dense(hidden_states)) + input_tensor)


class BertQuantizedAttention(nn.Module):
    def __init__(self, config, position_embedding_type=None, quant_config: dict = None):
        super().__init__()
        self.self = BertQuantizedSelfAttention(config, position_embedding_type=position_embedding_type, quant_config=quant_config)
        self.output = BertQuantizedSelfOutput(config, quant_config=quant_config["output"])

    def forward(self, hidden_states, attention_mask=None, head_mask=None, output_attentions=False, key_mask=None):
        dense = self.output.dense
        if (self.self.fused_eligible(hidden_states, key_mask, head_mask, output_attentions) and dense.accepts_prequantized()
                and output_quantizable(dense.config, hidden_states.shape[-1], hidden_states.shape[1])):
            # the x-quantizer of attention.output.dense runs in the attention epilogue; dense reads the bf16 operand
            oq = self.self.fused_forward(hidden_states, key_mask, out_cfg=dense.config)
            out = dense.forward_prequantized(oq).view(hidden_states.shape)
            return (self.output.LayerNorm(self.output.dropout(out) + hidden_states),)
        self_outputs = self.self(hidden_states, attention_mask, head_mask, output_attentions, key_mask=key_mask)
        return (self.output(self_outputs[0], hidden_states),) + self_outputs[1:]


class BertQuantizedIntermediate(nn.Module):
    def __init__(self, config, quant_config: dict):
        super().__init__()
        qc = quant_config["dense"]
        self.dense = get_quantized_cls("linear", qc)(config.hidden_size, config.intermediate_size, config=qc)
        self.quant_config = quant_config
        self.intermediate_act_fn = ACT2FN[config.hidden_act] if isinstance(config.hidden_act, str) else config.hidden_act

    def forward(self, hidden_states):
        return self.intermediate_act_fn(self.dense(hidden_states))


class BertQuantizedOutput(nn.Module):
    def __init__(self, config, quant_config):
        super().__init__()
        qc = quant_config["dense"]
        self.dense = get_quantized_cls("linear", qc)(config.intermediate_size, config.hidden_size, config=qc)
        self.quant_config = quant_config
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)

    def forward(self, hidden_states, input_tensor):
        return self.LayerNorm(self.dropout(self.dense(hidden_states)) + input_tensor)


class BertQuantizedLayer(nn.Module):
    def __init__(self, config, layer_num: int):
        super().__init__()
        if getattr(config, "is_decoder", False) or getattr(config, "add_cross_attention", False):
            raise NotImplementedError("decoder / cross-attention BERT is not on the quantized path (reference quant_config_bert.py:27)")
        qc = config.quant_config[f"model_layer_{layer_num}"]
        self.attention = BertQuantizedAttention(config, quant_config=qc["attention"])
        self.intermediate = BertQuantizedIntermediate(config, quant_config=qc["intermediate"])
        self.output = BertQuantizedOutput(config, quant_config=qc["output"])

    def forward(self, hidden_states, attention_mask=None, head_mask=None, output_attentions=False, key_mask=None):
        attn = self.attention(hidden_states, attention_mask, head_mask, output_attentions, key_mask=key_mask)
        attention_output = attn[0]
        layer_output = self.output(self.intermediate(attention_output), attention_output)
        return (layer_output,) + attn[1:]


class BertQuantizedEncoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.layer = nn.ModuleList([BertQuantizedLayer(config, i) for i in range(config.num_hidden_layers)])

    def forward(self, hidden_states, attention_mask=None, head_mask=None, output_attentions=False, output_hidden_states=False,
                key_mask=None):
        all_h, all_a = (), ()
        for i, layer in enumerate(self.layer):
            if output_hidden_states:
                all_h += (hidden_states,)
            out = layer(hidden_states, attention_mask, head_mask[i] if head_mask is not None else None, output_attentions,
                        key_mask=key_mask)
            hidden_states = out[0]
            if output_attentions:
                all_a += (out[1],)
        if output_hidden_states:
            all_h += (hidden_states,)
        return hidden_states, (all_h if output_hidden_states else None), (all_a if output_attentions else None)


class BertPooler(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.activation = nn.Tanh()

    def forward(self, hidden_states):
        return self.activation(self.dense(hidden_states[:, 0]))


class BertPredictionHeadTransform(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.transform_act_fn = ACT2FN[config.hidden_act] if isinstance(config.hidden_act, str) else config.hidden_act
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)

    def forward(self, hidden_states):
        return self.LayerNorm(self.transform_act_fn(self.dense(hidden_states)))


class BertLMPredictionHead(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.transform = BertPredictionHeadTransform(config)
        self.decoder = nn.Linear(config.hidden_size, config.vocab_size, bias=False)
        self.bias = nn.Parameter(torch.zeros(config.vocab_size))
        self.decoder.bias = self.bias

    def forward(self, hidden_states):
        return self.decoder(self.transform(hidden_states))


class BertOnlyMLMHead(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.predictions = BertLMPredictionHead(config)

    def forward(self, sequence_output):
        return self.predictions(sequence_output)


class BertQuantizedPreTrainedModel(PreTrainedModel):
    config_class = BertQuantizedConfig
    config: BertQuantizedConfig
    base_model_prefix = "bert"
    supports_gradient_checkpointing = False

    @torch.no_grad()
    def _init_weights(self, module):
        """reference :888-902"""
        std = self.config.initializer_range
        if isinstance(module, nn.Linear):
            module.weight.normal_(mean=0.0, std=std)
            if module.bias is not None:
                module.bias.zero_()
        elif isinstance(module, nn.Embedding):
            module.weight.normal_(mean=0.0, std=std)
            if module.padding_idx is not None:
                module.weight[module.padding_idx].zero_()
        elif isinstance(module, nn.LayerNorm):
            module.bias.zero_()
            module.weight.fill_(1.0)
        elif isinstance(module, BertLMPredictionHead):
            module.bias.zero_()


def _classification_loss(config, logits, labels, num_labels):
    """HF problem-type dispatch (reference :1815-1838).  The reference casts `labels` to the logits dtype BEFORE looking
    at their dtype (:1817), which makes the single-label branch unreachable unless `problem_type` is preset; here the
    dtype is inspected first (upstream HF behaviour) so integer class labels give cross-entropy."""
    if config.problem_type is None:
        if num_labels == 1:
            config.problem_type = "regression"
        elif num_labels > 1 and labels.dtype in (torch.long, torch.int):
            config.problem_type = "single_label_classification"
        else:
            config.problem_type = "multi_label_classification"
    if config.problem_type == "regression":
        labels = labels.to(logits.dtype)
        return MSELoss()(logits.squeeze(), labels.squeeze()) if num_labels == 1 else MSELoss()(logits, labels)
    if config.problem_type == "single_label_classification":
        return CrossEntropyLoss()(logits.view(-1, num_labels), labels.view(-1).long())
    return BCEWithLogitsLoss()(logits, labels.to(logits.dtype))


class BertQuantizedModel(BertQuantizedPreTrainedModel):
    def __init__(self, config, add_pooling_layer=True):
        super().__init__(config)
        self.config = config
        self.embeddings = BertEmbeddings(config)
        self.encoder = BertQuantizedEncoder(config)
        self.pooler = BertPooler(config) if add_pooling_layer else None
        self.fused_attention = True      # set False to force the op-by-op attention path (QUANTIZED_FUNC_MAP matmul functions)
        self.post_init()

    def get_input_embeddings(self):
        return self.embeddings.word_embeddings

    def set_input_embeddings(self, value):
        self.embeddings.word_embeddings = value

    def forward(self, input_ids=None, attention_mask=None, token_type_ids=None, position_ids=None, head_mask=None,
                inputs_embeds=None, output_attentions=None, output_hidden_states=None, return_dict=None):
        if input_ids is not None and inputs_embeds is not None:
            raise ValueError("You cannot specify both input_ids and inputs_embeds at the same time")
        if input_ids is None and inputs_embeds is None:
            raise ValueError("You have to specify either input_ids or inputs_embeds")
        shape = input_ids.size() if input_ids is not None else inputs_embeds.size()[:-1]
        device = input_ids.device if input_ids is not None else inputs_embeds.device
        if attention_mask is None:
            attention_mask = torch.ones(shape, device=device)
        dtype = self.embeddings.word_embeddings.weight.dtype
        # additive mask [B,1,1,S]: 0 where attended, finfo.min where padded (HF get_extended_attention_mask)
        ext = (1.0 - attention_mask[:, None, None, :].to(dtype)) * torch.finfo(dtype).min
        if head_mask is not None:
            if head_mask.dim() == 1:
                head_mask = head_mask[None, None, :, None, None].expand(self.config.num_hidden_layers, -1, -1, -1, -1)
            elif head_mask.dim() == 2:
                head_mask = head_mask[:, None, :, None, None]
        emb = self.embeddings(input_ids=input_ids, token_type_ids=token_type_ids, position_ids=position_ids,
                              inputs_embeds=inputs_embeds)
        # fused attention (bq_attention_masked) takes the 2-D 0/1 padding mask as a bitmap; every sequence must keep one key
        # (an all-masked row is a uniform distribution over the padding in the reference — op-by-op path)
        key_mask = None
        if self.fused_attention and emb.is_cuda and attention_mask.ndim == 2 and not torch.is_grad_enabled():
            valid = attention_mask != 0
            binary = bool(((attention_mask == 0) | (attention_mask == 1)).all())
            if binary and bool(valid.any(dim=1).all()):
                key_mask = key_mask_bits(valid)
        seq_out, all_h, all_a = self.encoder(emb, attention_mask=ext, head_mask=head_mask,
                                             output_attentions=bool(output_attentions),
                                             output_hidden_states=bool(output_hidden_states), key_mask=key_mask)
        pooled = self.pooler(seq_out) if self.pooler is not None else None
        if return_dict is False:
            return (seq_out, pooled) + tuple(v for v in (all_h, all_a) if v is not None)
        return BaseModelOutputWithPoolingAndCrossAttentions(last_hidden_state=seq_out, pooler_output=pooled,
                                                            hidden_states=all_h, attentions=all_a)


class BertQuantizedForSequenceClassification(BertQuantizedPreTrainedModel):
    """reference :1747-1850 — the class MODEL_MAP["bert"]["cls"] names."""

    def __init__(self, config):
        super().__init__(config)
        self.num_labels = config.num_labels
        self.config = config
        self.bert = BertQuantizedModel(config)
        p = config.classifier_dropout if config.classifier_dropout is not None else config.hidden_dropout_prob
        self.dropout = nn.Dropout(p)
        self.classifier = nn.Linear(config.hidden_size, config.num_labels)
        self.post_init()

    def forward(self, input_ids=None, attention_mask=None, token_type_ids=None, position_ids=None, head_mask=None,
                inputs_embeds=None, labels=None, output_attentions=None, output_hidden_states=None, return_dict=None):
        out = self.bert(input_ids, attention_mask=attention_mask, token_type_ids=token_type_ids, position_ids=position_ids,
                        head_mask=head_mask, inputs_embeds=inputs_embeds, output_attentions=output_attentions,
                        output_hidden_states=output_hidden_states)
        logits = self.classifier(self.dropout(out.pooler_output))
        loss = _classification_loss(self.config, logits, labels, self.num_labels) if labels is not None else None
        if return_dict is False:
            rest = (logits,) + tuple(v for v in (out.hidden_states, out.attentions) if v is not None)
            return ((loss,) + rest) if loss is not None else rest
        return SequenceClassifierOutput(loss=loss, logits=logits, hidden_states=out.hidden_states, attentions=out.attentions)


class BertQuantizedForMaskedLM(BertQuantizedPreTrainedModel):
    """reference :1504-1620"""
    _tied_weights_keys = {"cls.predictions.decoder.weight": "bert.embeddings.word_embeddings.weight",
                          "cls.predictions.decoder.bias": "cls.predictions.bias"}

    def __init__(self, config):
        super().__init__(config)
        self.bert = BertQuantizedModel(config, add_pooling_layer=False)
        self.cls = BertOnlyMLMHead(config)
        self.post_init()

    def get_output_embeddings(self):
        return self.cls.predictions.decoder

    def set_output_embeddings(self, new_embeddings):
        self.cls.predictions.decoder = new_embeddings
        self.cls.predictions.bias = new_embeddings.bias

    def forward(self, input_ids=None, attention_mask=None, token_type_ids=None, position_ids=None, head_mask=None,
                inputs_embeds=None, labels=None, output_attentions=None, output_hidden_states=None, return_dict=None):
        out = self.bert(input_ids, attention_mask=attention_mask, token_type_ids=token_type_ids, position_ids=position_ids,
                        head_mask=head_mask, inputs_embeds=inputs_embeds, output_attentions=output_attentions,
                        output_hidden_states=output_hidden_states)
        scores = self.cls(out.last_hidden_state)
        loss = None
        if labels is not None:
            loss = CrossEntropyLoss()(scores.view(-1, self.config.vocab_size), labels.view(-1))   # -100 = ignore
        if return_dict is False:
            rest = (scores,) + tuple(v for v in (out.hidden_states, out.attentions) if v is not None)
            return ((loss,) + rest) if loss is not None else rest
        return MaskedLMOutput(loss=loss, logits=scores, hidden_states=out.hidden_states, attentions=out.attentions)


class BertQuantizedForTokenClassification(BertQuantizedPreTrainedModel):
    """reference :1975-2060"""

    def __init__(self, config):
        super().__init__(config)
        self.num_labels = config.num_labels
        self.bert = BertQuantizedModel(config, add_pooling_layer=False)
        p = config.classifier_dropout if config.classifier_dropout is not None else config.hidden_dropout_prob
        self.dropout = nn.Dropout(p)
        self.classifier = nn.Linear(config.hidden_size, config.num_labels)
        self.post_init()

    def forward(self, input_ids=None, attention_mask=None, token_type_ids=None, position_ids=None, head_mask=None,
                inputs_embeds=None, labels=None, output_attentions=None, output_hidden_states=None, return_dict=None):
        out = self.bert(input_ids, attention_mask=attention_mask, token_type_ids=token_type_ids, position_ids=position_ids,
                        head_mask=head_mask, inputs_embeds=inputs_embeds, output_attentions=output_attentions,
                        output_hidden_states=output_hidden_states)
        logits = self.classifier(self.dropout(out.last_hidden_state))
        loss = CrossEntropyLoss()(logits.view(-1, self.num_labels), labels.view(-1)) if labels is not None else None
        if return_dict is False:
            rest = (logits,) + tuple(v for v in (out.hidden_states, out.attentions) if v is not None)
            return ((loss,) + rest) if loss is not None else rest
        return TokenClassifierOutput(loss=loss, logits=logits, hidden_states=out.hidden_states, attentions=out.attentions)


class BertQuantizedForQuestionAnswering(BertQuantizedPreTrainedModel):
    """reference :2064-2161"""

    def __init__(self, config):
        super().__init__(config)
        self.num_labels = config.num_labels
        self.bert = BertQuantizedModel(config, add_pooling_layer=False)
        self.qa_outputs = nn.Linear(config.hidden_size, config.num_labels)
        self.post_init()

    def forward(self, input_ids=None, attention_mask=None, token_type_ids=None, position_ids=None, head_mask=None,
                inputs_embeds=None, start_positions=None, end_positions=None, output_attentions=None,
                output_hidden_states=None, return_dict=None):
        out = self.bert(input_ids, attention_mask=attention_mask, token_type_ids=token_type_ids, position_ids=position_ids,
                        head_mask=head_mask, inputs_embeds=inputs_embeds, output_attentions=output_attentions,
                        output_hidden_states=output_hidden_states)
        start_logits, end_logits = self.qa_outputs(out.last_hidden_state).split(1, dim=-1)
        start_logits, end_logits = start_logits.squeeze(-1).contiguous(), end_logits.squeeze(-1).contiguous()
        loss = None
        if start_positions is not None and end_positions is not None:
            if start_positions.dim() > 1:
                start_positions = start_positions.squeeze(-1)
            if end_positions.dim() > 1:
                end_positions = end_positions.squeeze(-1)
            ignored = start_logits.size(1)
            start_positions, end_positions = start_positions.clamp(0, ignored), end_positions.clamp(0, ignored)
            ce = CrossEntropyLoss(ignore_index=ignored)
            loss = (ce(start_logits, start_positions) + ce(end_logits, end_positions)) / 2
        if return_dict is False:
            rest = (start_logits, end_logits) + tuple(v for v in (out.hidden_states, out.attentions) if v is not None)
            return ((loss,) + rest) if loss is not None else rest
        return QuestionAnsweringModelOutput(loss=loss, start_logits=start_logits, end_logits=end_logits,
                                            hidden_states=out.hidden_states, attentions=out.attentions)
