set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_consumers.py tests/test_gpu_models.py tests/test_gpu_fused_glue.py -x -q -k "not opt125m" > gpurun_out/s9_sanitizer_rest.log 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|misaligned" gpurun_out/s9_sanitizer_rest.log | head -8
