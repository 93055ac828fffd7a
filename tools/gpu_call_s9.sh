set -x
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_fused_glue.py -x -q -k "graphed" 2>&1 | tail -12
