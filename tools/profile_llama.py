"""Per-kernel time breakdown of a Llama-7B-shape W4A4 forward (torch.profiler, CUDA activities).
usage: python tools/profile_llama.py [layers=4] [block_minifloat|block_log] [batch=2]"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from torch.profiler import profile, ProfilerActivity
from llm_mixed_q_b200.models.llama_quantized import LlamaQuantizedConfig, LlamaQuantizedForCausalLM
layers = int(sys.argv[1]) if len(sys.argv) > 1 else 4
kind = sys.argv[2] if len(sys.argv) > 2 else "block_minifloat"
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda:0")
cfg = LlamaQuantizedConfig(quant_config=os.path.join(ROOT, "configs", f"llama_w4a4_{kind}.toml"), num_hidden_layers=layers,
                           initializer_range=1.28 if kind == "block_minifloat" else 0.02)
torch.manual_seed(0)
with torch.device(dev):
    model = LlamaQuantizedForCausalLM(cfg).eval()
ids = torch.randint(0, 32000, (batch, 2048), device=dev)
with torch.no_grad():
    for _ in range(2): model(input_ids=ids, labels=ids)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        model(input_ids=ids, labels=ids)
        torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    t = getattr(e, "device_time_total", None)
    if t is None: t = getattr(e, "cuda_time_total", 0)
    if t > 0 and e.device_type == torch.autograd.DeviceType.CUDA:
        rows.append((t / 1e3, e.count, e.key[:120]))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(f"layers {layers} kind {kind} batch {batch}: total device ms {tot:.2f}")
for r in rows[:40]:
    print(f"{r[0]:9.3f} ms  {100*r[0]/tot:5.1f}%  x{r[1]:<5d} {r[2]}")
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open(f"gpurun_out/profile_llama_{kind}.json", "w"))
