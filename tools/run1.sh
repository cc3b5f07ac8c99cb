set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | head -2; nproc; free -g | head -2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02_pytest1.log
tail -5 gpurun_out/r02_pytest1.log
CLOOPS_BENCH_VERBOSE=1 timeout 600 python bench.py --config 4 --pets 20000000 --steps 2 --warmup 3 > gpurun_out/r02_bench_c4_20M.json 2> gpurun_out/r02_bench_c4_20M.err
tail -c 3000 gpurun_out/r02_bench_c4_20M.json; tail -5 gpurun_out/r02_bench_c4_20M.err
timeout 900 python bench.py --config 4 --steps 3 --warmup 3 > gpurun_out/r02_bench_c4.json 2> gpurun_out/r02_bench_c4.err
tail -c 3000 gpurun_out/r02_bench_c4.json; tail -5 gpurun_out/r02_bench_c4.err
