set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/s9_gpu_tests.log; tail -3 gpurun_out/s9_gpu_tests.log
timeout 300 python tools/bench_kernels.py fused 2>&1 | tail -22
