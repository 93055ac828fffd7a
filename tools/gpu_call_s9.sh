# session-9 GPU call: LN kernel tests + micro-bench, GEMM tests, bench, ncu --set full of one layer (raw CSV only)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_fused_glue.py tests/test_gpu_consumers.py -x -q 2>&1 | tail -5
timeout 300 python tools/bench_kernels.py fused 2>&1 | head -12
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/s9_bench.json 2> gpurun_out/s9_bench.err
tail -c 300 gpurun_out/s9_bench.json
timeout 600 ncu --set full --clock-control none --profile-from-start off -f -o /tmp/s9_layer python tools/ncu_layer.py > gpurun_out/s9_ncu_layer.log 2>&1
tail -2 gpurun_out/s9_ncu_layer.log
ncu -i /tmp/s9_layer.ncu-rep --page raw --csv > gpurun_out/s9_layer_raw.csv 2>/dev/null
du -sh gpurun_out
