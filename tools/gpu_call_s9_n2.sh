# BASELINE configs[3] at N GPUs: Llama-7B W4A4 block_minifloat + block_log, data-parallel replicas
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29588 tools/bench_configs.py --config 4 --format both > gpurun_out/s9_cfg4_n$N.log 2>&1
grep -E '^\{"metric"' gpurun_out/s9_cfg4_n$N.log | cut -c1-330
tail -3 gpurun_out/s9_cfg4_n$N.log | cut -c1-300
