# session-9 validation recipe (one B200): full GPU tests, smoke, bench (+ reference arm), ncu launch list, ncu --set full of one layer.
#   gpurun --timeout 1800 -- 'bash tools/gpu_call_s9.sh'
# Outputs land in gpurun_out/ (kept under 64 MiB: the .ncu-rep stays on the box, only its raw CSV page comes back).
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/s9_gpu_tests.log; tail -2 gpurun_out/s9_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/s9_bench.json 2> gpurun_out/s9_bench.err; tail -3 gpurun_out/s9_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s9_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/s9_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sub --eager-e2e > gpurun_out/s9_bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --profile-from-start off -f -o /tmp/s9_layer python tools/ncu_layer.py > gpurun_out/s9_ncu_layer.log 2>&1
ncu -i /tmp/s9_layer.ncu-rep --page raw --csv > gpurun_out/s9_layer_raw.csv 2>/dev/null
timeout 300 python tools/bench_kernels.py fused > gpurun_out/s9_bench_kernels_fused.log 2>&1
du -sh gpurun_out
