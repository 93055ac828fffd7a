# session-9 GPU call: batched split GEMM (general-route matmul) tests + Llama block_log timing
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/s9_gpu_tests.log; tail -8 gpurun_out/s9_gpu_tests.log
timeout 300 python tools/profile_llama.py 2 block_log 2 > gpurun_out/s9_profile_llama_block_log.txt 2>&1; head -16 gpurun_out/s9_profile_llama_block_log.txt | cut -c1-150
timeout 600 python tools/bench_configs.py --config 4 --format block_log > gpurun_out/s9_cfg4_bl.log 2>&1; tail -2 gpurun_out/s9_cfg4_bl.log | cut -c1-400
