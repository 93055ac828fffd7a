# session-9 two-GPU call: peer tests on two real devices, config-5 column-parallel at 2 GPUs, LN kernel check
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_peer.py tests/test_gpu_fused_glue.py -x -q 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/bench_configs.py --config 5 > gpurun_out/s9_cfg5_n2.log 2>&1
grep -E "layer total|fc1" gpurun_out/s9_cfg5_n2.log | cut -c1-900
timeout 300 python tools/bench_kernels.py fused 2>&1 | head -12
