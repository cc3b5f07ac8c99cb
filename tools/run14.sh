set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dbscan.py tests/test_gpu_fuzz.py tests/test_gpu_edge.py -x -q 2>&1 | tail -4
timeout 400 python tools/count_probe.py > gpurun_out/r02_count_probe_8probes.log 2>&1
cat gpurun_out/r02_count_probe_8probes.log
