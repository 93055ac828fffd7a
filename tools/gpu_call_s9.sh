set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused_glue.py tests/test_gpu_models.py -x -q 2>&1 | tail -6
timeout 300 python tools/profile_llama.py 4 block_minifloat 2 > gpurun_out/s9_profile_llama_bmf.txt 2>&1; head -14 gpurun_out/s9_profile_llama_bmf.txt | cut -c1-160
timeout 600 python tools/bench_configs.py --config 4 --format block_minifloat > gpurun_out/s9_cfg4_bmf.log 2>&1; tail -1 gpurun_out/s9_cfg4_bmf.log | cut -c1-300
