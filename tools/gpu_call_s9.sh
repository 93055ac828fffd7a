set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/s9_gpu_tests.log; tail -2 gpurun_out/s9_gpu_tests.log
timeout 300 python tools/bench_kernels.py fused 2>&1 | head -11
timeout 300 python tools/profile_llama.py 4 block_minifloat 2 > gpurun_out/s9_profile_llama_bmf.txt 2>&1; head -12 gpurun_out/s9_profile_llama_bmf.txt | cut -c1-160
timeout 600 python tools/bench_configs.py --config 4 --format block_minifloat > gpurun_out/s9_cfg4_bmf.log 2>&1; tail -1 gpurun_out/s9_cfg4_bmf.log | cut -c1-300
