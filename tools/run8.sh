set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 600 python tools/count_probe.py > gpurun_out/r02_count_probe_final.log 2>&1
cat gpurun_out/r02_count_probe_final.log
for mode in ranked legacy; do
CLOOPS_RC=$mode timeout 900 python bench.py --config 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c4_$mode.json 2> gpurun_out/r02_bench_c4_$mode.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_c4_$mode.json').read().strip().splitlines()[-1])
print("$mode", {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline_range_count'])
print(d['stages_ms'])
PY
done
CLOOPS_RC=ranked timeout 900 python -m pytest tests/test_gpu_scoring.py tests/test_gpu_fullsize.py tests/test_gpu_pipe.py tests/test_gpu_scripts.py -x -q 2>&1 | tail -4
