# session-9 GPU call: epilogue changes — tests, micro-bench, bench
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/s9_gpu_tests.log; tail -4 gpurun_out/s9_gpu_tests.log
timeout 300 python tools/bench_kernels.py fused 2>&1 | tail -8
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/s9_bench.json 2> gpurun_out/s9_bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/s9_bench.json'))
print(d['value'], d['ms_per_step'], d['clocks'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['share_of_step'])
P
