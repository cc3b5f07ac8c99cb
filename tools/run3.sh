set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_dbscan.py tests/test_gpu_fuzz.py tests/test_gpu_edge.py tests/test_gpu_fullsize.py tests/test_gpu_pipe.py -x -q 2>&1 | tail -8 > gpurun_out/r02_pytest3.log
tail -8 gpurun_out/r02_pytest3.log
timeout 900 python tools/step_profile.py > gpurun_out/r02_step_profile.log 2>&1
grep -v "^$" gpurun_out/r02_step_profile.log | head -80
timeout 600 python tools/pass_probe.py 0 2>&1 | grep -v "^\[cloops\]" | tail -12
