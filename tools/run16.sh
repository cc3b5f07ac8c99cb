set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for t in 512 128; do
CLOOPS_RC_TEAM=$t timeout 600 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/r02_bench_c4_team$t.json 2> /dev/null
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_c4_team$t.json').read().strip().splitlines()[-1])
print("team $t", {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['ms_per_step'], d['roofline_range_count']['ms'], d['stages_ms']['range_counts'])
PY
done
