set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for n in 8 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2956$n bench.py --gpus $n > gpurun_out/r02_bench_c4_n$n.json 2> gpurun_out/r02_bench_c4_n$n.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_c4_n$n.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches','lpt_ideal_speedup','rank_balance')}, d['e2e']['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline_range_count']['frac'])
PY
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 tools/dist_loops_check.py 6000000 4 2>&1 | grep "IDENTICAL\|rror" | tail -3
