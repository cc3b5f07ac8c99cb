set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scoring.py tests/test_gpu_fullsize.py tests/test_gpu_pipe.py tests/test_gpu_scripts.py tests/test_gpu_rounds.py -x -q 2>&1 | tail -4
for mode in team warp; do
CLOOPS_RC=$mode timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_c4_$mode.json 2> gpurun_out/r02_bench_c4_$mode.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_c4_$mode.json').read().strip().splitlines()[-1])
print("$mode", {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline_range_count']['ms'], d['roofline_range_count']['frac'])
print(d['stages_ms'])
PY
done
