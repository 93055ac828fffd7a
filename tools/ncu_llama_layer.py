"""One steady-state forward of a 1-layer Llama-7B-shaped W4A4 block_minifloat model (batch 2 x 2048) inside a cudaProfiler range: the
kernels of the fused Llama layer — RMSNorm+quantize, q / k GEMMs with the RoPE epilogue, v GEMM, attention (head_dim 128), o_proj, the
gate||up GEMM with the gated-SiLU epilogue, down_proj — captured once each.

usage (one GPU): ncu --set full --clock-control none --profile-from-start off -f -o /tmp/llama_layer python tools/ncu_llama_layer.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llm_mixed_q_b200.models.llama_quantized import LlamaQuantizedConfig, LlamaQuantizedForCausalLM   # noqa: E402
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
kind = sys.argv[1] if len(sys.argv) > 1 else "block_minifloat"
dev = torch.device("cuda:0")
cfg = LlamaQuantizedConfig(quant_config=os.path.join(ROOT, "configs", f"llama_w4a4_{kind}.toml"), num_hidden_layers=1,
                           initializer_range=1.28 if kind == "block_minifloat" else 0.02)
torch.manual_seed(0)
with torch.device(dev):
    model = LlamaQuantizedForCausalLM(cfg).eval()
ids = torch.randint(0, 32000, (2, 2048), device=dev)
with torch.no_grad():
    for _ in range(2):
        model(input_ids=ids, labels=ids)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    out = model(input_ids=ids, labels=ids)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("loss", float(out.loss))
