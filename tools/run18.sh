set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for t in 3 4; do
CLOOPS_STREAMS=$t timeout 600 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/r02_bench_c4_streams$t.json 2> gpurun_out/r02_bench_c4_streams$t.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_c4_streams$t.json').read().strip().splitlines()[-1])
print("streams $t", {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['ms_per_step'], d['result'], d['roofline']['frac'])
PY
tail -2 gpurun_out/r02_bench_c4_streams$t.err
done
