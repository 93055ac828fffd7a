# session-9 final validation: full GPU tests, smoke, bench, ncu launch list of the bench command, ncu --set full of one layer
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/s9_gpu_tests.log; tail -2 gpurun_out/s9_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/s9_bench.json 2> gpurun_out/s9_bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/s9_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['share_of_step'], d['gpu_launches'])
P
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/s9_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sub > gpurun_out/s9_bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --profile-from-start off -f -o /tmp/s9_layer python tools/ncu_layer.py > gpurun_out/s9_ncu_layer.log 2>&1
ncu -i /tmp/s9_layer.ncu-rep --page raw --csv > gpurun_out/s9_layer_raw.csv 2>/dev/null
timeout 300 python tools/bench_kernels.py fused > gpurun_out/s9_bench_kernels_fused.log 2>&1
du -sh gpurun_out
