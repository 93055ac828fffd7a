set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none -k regex:norm_quant -c 3 -f -o /tmp/ln python tools/ncu_micro.py > gpurun_out/s9_ncu_ln.log 2>&1
tail -3 gpurun_out/s9_ncu_ln.log
ncu -i /tmp/ln.ncu-rep --page raw --csv > gpurun_out/s9_ln_raw.csv 2>/dev/null
