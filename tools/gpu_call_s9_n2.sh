# session-9 two-GPU call: peer tests on two real devices, bench.py under torchrun (N=2), config-5 column-parallel at 2 GPUs
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
timeout 300 python -m pytest tests/test_gpu_peer.py -x -q 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/s9_bench_n2.json 2> gpurun_out/s9_bench_n2.err
tail -c 400 gpurun_out/s9_bench_n2.json; tail -5 gpurun_out/s9_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/bench_configs.py --config 5 > gpurun_out/s9_cfg5_n2.log 2>&1
tail -8 gpurun_out/s9_cfg5_n2.log | cut -c1-600
timeout 300 python bench.py --impl reference --gpus 2 --steps 2 --warmup 1 | cut -c1-300
