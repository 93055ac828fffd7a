# round-2 validation recipe (one B200): full GPU tests, smoke, bench (+ reference arm), ncu launch list, ncu --set full of one layer and of the
# split-attention kernels, kernel micro-benchmarks, secondary configs.
#   gpurun --timeout 2400 -- 'bash tools/gpu_call_r02.sh'
# Outputs land in gpurun_out/r02_* (kept under 64 MiB: the .ncu-rep files stay on the box, only their raw CSV pages come back).
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_gpu_tests.log; tail -2 gpurun_out/r02_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; tail -3 gpurun_out/r02_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sub --eager-e2e > gpurun_out/r02_bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --profile-from-start off -f -o /tmp/r02_layer python tools/ncu_layer.py > gpurun_out/r02_ncu_layer.log 2>&1
ncu -i /tmp/r02_layer.ncu-rep --page raw --csv > gpurun_out/r02_layer_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none -k regex:"softmax_quant|gemm_bf16_tn|rope_split|split3_transposed" -c 8 -f -o /tmp/r02_split python tools/run_split_attention_once.py 4 > gpurun_out/r02_ncu_split.log 2>&1
ncu -i /tmp/r02_split.ncu-rep --page raw --csv > gpurun_out/r02_split_raw.csv 2>/dev/null
timeout 300 python tools/bench_kernels.py quant > gpurun_out/r02_bench_kernels_stream.log 2>&1
timeout 300 python tools/bench_kernels.py fused > gpurun_out/r02_bench_kernels_fused.log 2>&1
timeout 300 python tools/bench_xform.py > gpurun_out/r02_bench_xform.log 2>&1
timeout 300 python tools/bench_attention.py > gpurun_out/r02_bench_attention.log 2>&1
timeout 300 python tools/bench_attention.py d128 > gpurun_out/r02_bench_attention_d128.log 2>&1
timeout 300 python tools/run_split_attention_once.py 4 time > gpurun_out/r02_split_attention_times.json 2>&1
rm -f gpurun_out/bench_configs.jsonl
timeout 600 python tools/bench_configs.py --config 4 --format both --batch 2 --steps 5 --warmup 2 --graph > /dev/null 2>&1
timeout 300 python tools/bench_configs.py --config 0 --batch 32 --steps 10 --warmup 3 > /dev/null 2>&1
cp gpurun_out/bench_configs.jsonl gpurun_out/r02_bench_configs.jsonl
timeout 300 python tools/profile_llama.py 2 block_log 2 > gpurun_out/r02_profile_llama_block_log.txt 2>&1
timeout 300 python tools/profile_llama.py 2 block_minifloat 2 > gpurun_out/r02_profile_llama_bmf.txt 2>&1
du -sh gpurun_out
