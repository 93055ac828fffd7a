# multi-GPU: config-5 column-parallel layer at N GPUs (N = number of visible devices) + peer tests
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 tools/bench_configs.py --config 5 > gpurun_out/s9_cfg5_n$N.log 2>&1
grep -E "layer total|fc1|Error|error" gpurun_out/s9_cfg5_n$N.log | cut -c1-1000
