# session-9 final validation: full GPU tests, smoke, bench (+ reference arm)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/s9_gpu_tests.log; tail -2 gpurun_out/s9_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/s9_bench.json 2> gpurun_out/s9_bench.err; tail -3 gpurun_out/s9_bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/s9_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e'], d['clocks'], d['roofline']['achieved'], d['roofline']['frac'], d['gpu_launches'])
P
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s9_bench_ref.json; cut -c1-160 gpurun_out/s9_bench_ref.json
