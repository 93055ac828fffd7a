"""Per-kernel time breakdown of one OPT-1.3B W6A6 forward (torch.profiler, CUDA activities)."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda:0")
layers = int(sys.argv[1]) if len(sys.argv) > 1 else None
model = bench.build_model(dev, layers=layers)
ids = torch.randint(0, 50272, (bench.BATCH, bench.SEQ), device=dev)
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
        rows.append((t / 1e3, e.count, e.key[:110]))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(f"total device ms {tot:.1f}")
for r in rows[:40]:
    print(f"{r[0]:9.2f} ms  {100*r[0]/tot:5.1f}%  x{r[1]:<5d} {r[2]}")
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/profile_step.json", "w"))
