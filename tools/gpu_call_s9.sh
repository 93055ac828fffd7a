set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_models.py -x -q -s -k "1p3b" > gpurun_out/s9_opt13b_parity.log 2>&1
grep -n "opt-1.3b fused\|passed\|failed\|Error" gpurun_out/s9_opt13b_parity.log | cut -c1-400
