# multi-GPU: bench.py under torchrun at N GPUs (N = number of visible devices)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/s9_bench_n$N.json 2> gpurun_out/s9_bench_n$N.err
python - <<P
import json
d=json.loads(open('gpurun_out/s9_bench_n$N.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['mode'], d['roofline']['frac'])
P
tail -3 gpurun_out/s9_bench_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29578 bench.py --impl reference --gpus $N --steps 1 --warmup 0 | cut -c1-200
