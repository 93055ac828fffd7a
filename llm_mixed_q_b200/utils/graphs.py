"""
CUDA-graph replay of a quantized causal-LM forward (PTQ inference, fixed batch x sequence shape).

One forward of OPT-1.3B is ~230 kernel launches issued from Python through ctypes; a slow or busy host core shows up directly in
the step time whenever the launch queue runs dry (every step that reads the loss back does that).  `GraphedForward` captures the
whole forward once — our kernels are plain stream launches, so they are captured like any torch op — and replays it per step:
host cost per step is one H2D copy of the token ids, one graph launch and the read-back of the loss.

    runner = GraphedForward(model, batch=8, seq_len=2048)      # runs warm-up forwards (PTQ weight overwrite, caches), then captures
    loss = runner(ids)                                          # ids: int64 [batch, seq_len], host (pinned) or device
    runner.logits                                               # static output buffers, overwritten by every replay

Falls back to the eager forward (same results) if the capture fails, e.g. for an input that needs a host-side decision inside the
forward (`runner.graph is None` then, and `runner.error` says why).
"""
from __future__ import annotations

import torch


class GraphedForward:
    def __init__(self, model, batch: int, seq_len: int, device=None, warmup: int = 2, with_labels: bool = True):
        self.model = model
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        self.with_labels = with_labels
        self.ids = torch.zeros((batch, seq_len), dtype=torch.int64, device=self.device)
        self.graph, self.error = None, None
        self.loss = self.logits = None
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(warmup, 1)):
                self._forward()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        try:
            g = torch.cuda.CUDAGraph()
            with torch.no_grad(), torch.cuda.graph(g):
                out = self._forward()
            self.graph, self.loss, self.logits = g, out.loss, out.logits
        except Exception as e:                       # not capturable: keep working, eagerly
            self.error = f"{type(e).__name__}: {e}"
            torch.cuda.synchronize(self.device)

    def _forward(self):
        return self.model(input_ids=self.ids, labels=self.ids if self.with_labels else None)

    @torch.no_grad()
    def __call__(self, input_ids: torch.Tensor):
        self.ids.copy_(input_ids, non_blocking=True)
        if self.graph is not None:
            self.graph.replay()
        else:
            out = self._forward()
            self.loss, self.logits = out.loss, out.logits
        return self.loss
