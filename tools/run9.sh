set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_loops_check.py 6000000 4 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -6
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_c4_n2.json 2> gpurun_out/r02_bench_c4_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_c4_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches','lpt_ideal_speedup','rank_balance')}, d['e2e'], d['roofline']['frac'])
print(d['stages_ms'])
PY
tail -5 gpurun_out/r02_bench_c4_n2.err
