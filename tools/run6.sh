set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 600 python tools/pass_probe.py 0 2>&1 | grep -v "^\[cloops\]" | tail -10
CLOOPS_RQ=tiled timeout 900 python bench.py --config 4 --steps 3 --warmup 3 > gpurun_out/r02_bench_c4c.json 2> gpurun_out/r02_bench_c4c.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_c4c.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','result')}, d['e2e'], d['cpu_baseline'], d['roofline'], d['roofline_range_count'])
print(d['stages_ms'])
PY
tail -3 gpurun_out/r02_bench_c4c.err
