# end-of-round-2 validation recipe (one B200), state of HEAD after the Llama GEMM epilogues / attention tiers / PDL switch:
#   gpurun --timeout 2400 -- 'bash tools/gpu_call_r02_s5.sh'
# Outputs land in gpurun_out/r02_*_s5.* (kept under 64 MiB: .ncu-rep files stay on the box, only their raw CSV pages come back).
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_gpu_tests_s5.log; tail -2 gpurun_out/r02_gpu_tests_s5.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1_s5.json 2> gpurun_out/r02_bench_n1_s5.err; tail -3 gpurun_out/r02_bench_n1_s5.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref_s5.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_s5.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sub --eager-e2e > gpurun_out/r02_bench_under_ncu_s5.log 2>&1
timeout 600 ncu --set full --clock-control none --profile-from-start off -f -o /tmp/r02_layer python tools/ncu_layer.py > gpurun_out/r02_ncu_layer_s5.log 2>&1
ncu -i /tmp/r02_layer.ncu-rep --page raw --csv > gpurun_out/r02_layer_raw_s5.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --profile-from-start off -f -o /tmp/r02_llama python tools/ncu_llama_layer.py > gpurun_out/r02_ncu_llama_layer_s5.log 2>&1
ncu -i /tmp/r02_llama.ncu-rep --page raw --csv > gpurun_out/r02_llama_layer_raw_s5.csv 2>/dev/null
timeout 300 python tools/bench_attention.py > gpurun_out/r02_bench_attention_s5.log 2>&1
cp gpurun_out/bench_attention.json gpurun_out/r02_bench_attention_s5.json
timeout 300 python tools/bench_attention.py peaked > gpurun_out/r02_bench_attention_peaked_s5.log 2>&1
timeout 300 python tools/bench_attention.py d128 > gpurun_out/r02_bench_attention_d128_s5.log 2>&1
timeout 300 python tools/bench_llama_epilogues.py > gpurun_out/r02_bench_llama_epilogues_s5.log 2>&1
timeout 300 python tools/bench_kernels.py fused > gpurun_out/r02_bench_kernels_fused_s5.log 2>&1
rm -f gpurun_out/bench_configs.jsonl
timeout 600 python tools/bench_configs.py --config 4 --format both --batch 2 --steps 5 --warmup 2 --graph > /dev/null 2>&1
timeout 300 python tools/bench_configs.py --config 0 --batch 32 --steps 10 --warmup 3 > /dev/null 2>&1
cp gpurun_out/bench_configs.jsonl gpurun_out/r02_bench_configs_s5.jsonl
du -sh gpurun_out
