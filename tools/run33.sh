set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python bench.py > gpurun_out/r02_bench_c4_final2.json 2> gpurun_out/r02_bench_c4_final2.err
tail -c 1500 gpurun_out/r02_bench_c4_final2.json
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/r02_launches_step_final.csv python tools/ncu_step.py 4 200000000 3 2>&1 | tail -2
python tools/launch_shares2.py gpurun_out/r02_launches_step_final.csv gpurun_out/r02_step_kernel_shares_final.csv
head -25 gpurun_out/r02_step_kernel_shares_final.csv
