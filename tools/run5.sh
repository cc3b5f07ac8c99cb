set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ncu -k regex:count_kernel --set full --clock-control none --import-source on -f -o gpurun_out/r02_rq_split python tools/count_probe.py ncu > gpurun_out/ncu_split.log 2>&1
tail -3 gpurun_out/ncu_split.log
CLOOPS_RQ=tiled ncu -k regex:count_kernel --set full --clock-control none --import-source on -f -o gpurun_out/r02_rq_tiled python tools/count_probe.py ncu > gpurun_out/ncu_tiled.log 2>&1
tail -3 gpurun_out/ncu_tiled.log
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 600 python tools/pass_probe.py 0 2>&1 | grep -v "^\[cloops\]" | tail -10
