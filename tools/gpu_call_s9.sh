set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_fused_glue.py -x -q -k "graphed or fused_opt_layer" 2>&1 | tail -12
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sub > gpurun_out/s9_bench.json 2> gpurun_out/s9_bench.err; tail -5 gpurun_out/s9_bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/s9_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e'], d['clocks'])
P
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sub --eager-e2e > gpurun_out/s9_bench_eager.json 2>/dev/null
python - <<'P'
import json
d=json.load(open('gpurun_out/s9_bench_eager.json'))
print(d['value'], d['ms_per_step'], d['e2e'])
P
