set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/bench_kernels.py fused 2>&1 | grep attention
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/s9_bench.json 2> gpurun_out/s9_bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/s9_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['share_of_step'])
for k,v in d['roofline']['other_kernels'].items(): print(k, {a:b for a,b in v.items() if a!='note'})
P
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s9_bench_ref.json; cut -c1-300 gpurun_out/s9_bench_ref.json
