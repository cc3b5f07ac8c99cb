set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_dbscan.py tests/test_gpu_fuzz.py tests/test_gpu_edge.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -12
timeout 600 python tools/count_probe.py > gpurun_out/r02_count_probe_run.log 2>&1
cat gpurun_out/r02_count_probe_run.log
CLOOPS_RQ=tiled timeout 600 python tools/count_probe.py > gpurun_out/r02_count_probe_tiled.log 2>&1
cat gpurun_out/r02_count_probe_tiled.log
