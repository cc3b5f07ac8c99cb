set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nproc
for t in 2 1; do
CLOOPS_STREAMS=$t timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2954$t bench.py --gpus 8 --steps 3 > gpurun_out/r02_bench_c4_n8_s$t.json 2> /dev/null
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_c4_n8_s$t.json').read().strip().splitlines()[-1])
print("streams $t", {k:d[k] for k in ('value','ms_per_step','n_gpus','rank_balance')}, d['e2e']['ms_per_step'])
PY
done
