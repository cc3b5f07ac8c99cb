set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/dist_loops_check.py 6000000 4 2>&1 | grep "IDENTICAL\|Error\|error" | tail -4
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/r02_bench_c4_n$n.json 2> gpurun_out/r02_bench_c4_n$n.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_c4_n$n.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches','lpt_ideal_speedup','rank_balance')}, d['e2e']['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])
PY
done
