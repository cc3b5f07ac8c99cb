set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'dist_stats_kernel|hist_' -o gpurun_out/stats python tools/stats_target.py 2>&1 | tail -4
ls -la gpurun_out/
