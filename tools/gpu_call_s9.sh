# session-9 GPU call: LN kernel tests + micro-bench, full GPU tests, bench, ncu launch list, ncu --set full of one layer
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_fused_glue.py -x -q -k "layernorm" 2>&1 | tail -15
timeout 300 python tools/bench_kernels.py fused 2>&1 | tail -25
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/s9_gpu_tests.log
tail -3 gpurun_out/s9_gpu_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/s9_bench.json 2> gpurun_out/s9_bench.err
tail -c 300 gpurun_out/s9_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/s9_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sub > gpurun_out/s9_bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --profile-from-start off -f -o /tmp/s9_layer python tools/ncu_layer.py > gpurun_out/s9_ncu_layer.log 2>&1
tail -2 gpurun_out/s9_ncu_layer.log
ncu -i /tmp/s9_layer.ncu-rep --page raw --csv > gpurun_out/s9_layer_raw.csv 2>/dev/null
ls -la /tmp/s9_layer.ncu-rep gpurun_out/
sz=$(stat -c %s /tmp/s9_layer.ncu-rep); if [ "$sz" -lt 30000000 ]; then cp /tmp/s9_layer.ncu-rep gpurun_out/; fi
du -sh gpurun_out
