"""One steady-state forward of a 1-layer OPT-1.3B W6A6 model (same layer shapes as bench.py) inside a cudaProfiler range.

usage (one GPU): ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/layer \
                     python tools/ncu_layer.py
Every kernel of one layer + lm_head + loss is captured once; tools/ncu_extract.py turns the report into JSON."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
dev = torch.device("cuda:0")
model = bench.build_model(dev, layers=1)
ids = torch.randint(0, 50272, (bench.BATCH, bench.SEQ), device=dev)
with torch.no_grad():
    for _ in range(2):
        model(input_ids=ids, labels=ids)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    out = model(input_ids=ids, labels=ids)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("loss", float(out.loss))
