set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/r02_bench_c4_n2_final.json 2> gpurun_out/r02_bench_c4_n2_final.err
tail -c 700 gpurun_out/r02_bench_c4_n2_final.json
