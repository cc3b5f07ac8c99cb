set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_c4d.json 2> gpurun_out/r02_bench_c4d.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_c4d.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline_range_count']['ms'])
print(d['stages_ms'])
PY
timeout 300 python tools/pass_probe.py 0 2>&1 | grep -v "^\[cloops\]" | tail -10
