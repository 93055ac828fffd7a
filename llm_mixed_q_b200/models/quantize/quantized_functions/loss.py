"""
Shifted token cross-entropy of a causal LM (host side of bq_token_ce_mean, csrc/loss.cu).

The reference ends every causal-LM forward with
    shift_logits = logits[..., :-1, :].contiguous(); shift_labels = labels[..., 1:].contiguous()
    loss = CrossEntropyLoss()(shift_logits.view(-1, V), shift_labels.view(-1))
(models/opt_quantized/modeling_opt.py:1086-1098, models/llama_quantized/modeling_llama.py:867-879) and
eval/eval_lm.py:41-63 turns that loss into perplexity.  At OPT-1.3B, batch 8 x seq 2048, the logits are 3.3 GB: the copy
and the log-softmax move ~13 GB; the kernel reads the logits once.  Forward only (like the rest of this package).
"""
from __future__ import annotations

import torch

from .... import _lib as L


def causal_lm_loss(logits: torch.Tensor, labels: torch.Tensor, shift: bool = True, ignore_index: int = -100) -> torch.Tensor:
    """Mean cross-entropy of logits [B, S, V] (fp32, CUDA) against labels [B, S] (shifted by one when `shift`).
    Returns a 0-dim fp32 tensor, like CrossEntropyLoss()."""
    if torch.is_grad_enabled() and logits.requires_grad:
        # the kernel is forward-only: keep the reference's differentiable ops when a graph is being recorded
        lg, lb = (logits[..., :-1, :], labels[..., 1:]) if shift else (logits, labels)
        return torch.nn.functional.cross_entropy(lg.reshape(-1, lg.shape[-1]), lb.reshape(-1).to(lg.device),
                                                 ignore_index=ignore_index)
    L.require_cuda_f32(logits, "logits")
    if logits.dim() == 2:
        logits = logits.unsqueeze(0)
        labels = labels.unsqueeze(0)
    if logits.dim() != 3 or labels.shape != logits.shape[:2]:
        raise ValueError(f"logits {tuple(logits.shape)} / labels {tuple(labels.shape)}: expected [B, S, V] and [B, S]")
    B, S, V = logits.shape
    if logits.stride(2) != 1 or (B > 1 and logits.stride(0) != S * logits.stride(1)):
        logits = logits.contiguous()
    labels = labels.to(device=logits.device, dtype=torch.int64).contiguous()
    lib = L.load()
    ws_bytes = lib.bq_token_ce_workspace_bytes(B, S)
    ws = torch.empty(max(ws_bytes, 4), dtype=torch.uint8, device=logits.device)
    out = torch.empty(2, dtype=torch.float32, device=logits.device)
    L.check(lib.bq_token_ce_mean(logits.data_ptr(), B, S, V, logits.stride(1), labels.data_ptr(), 1 if shift else 0, ignore_index,
                                 out.data_ptr(), ws.data_ptr(), ws_bytes, L.stream_ptr(logits.device)), "bq_token_ce_mean")
    return out[0]
