set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py > gpurun_out/r02_bench_c4.json 2> gpurun_out/r02_bench_c4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_c4.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','result','clocks')}, d['e2e'], d['cpu_baseline'], d['roofline'], d['roofline_range_count'])
print(d['stages_ms'])
PY
wc -l gpurun_out/r02_bench_c4.json
timeout 900 python bench.py --impl reference > gpurun_out/r02_bench_c4_reference.json 2> gpurun_out/r02_bench_c4_reference.err
tail -c 900 gpurun_out/r02_bench_c4_reference.json
