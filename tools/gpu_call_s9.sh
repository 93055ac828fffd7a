set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_models.py -x -q -s -k opt125m > gpurun_out/s9_opt125m.log 2>&1
grep -n "AssertionError\|opt125m config-1\|passed\|failed" gpurun_out/s9_opt125m.log | cut -c1-600
