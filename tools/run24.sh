set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for r in 1 0; do
CLOOPS_REUSE_INDEX=$r timeout 600 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/r02_bench_c4_reuse$r.json 2> gpurun_out/r02_bench_c4_reuse$r.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_c4_reuse$r.json').read().strip().splitlines()[-1])
print("reuse $r", {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['ms_per_step'], d['result'], d['roofline']['frac'])
print(d['stages_ms'])
PY
tail -2 gpurun_out/r02_bench_c4_reuse$r.err
done
