set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 400 python tools/count_probe.py > gpurun_out/r02_count_probe_moff.log 2>&1
cat gpurun_out/r02_count_probe_moff.log
timeout 600 python bench.py > gpurun_out/r02_bench_c4.json 2> gpurun_out/r02_bench_c4.err
wc -l gpurun_out/r02_bench_c4.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_c4.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','result','clocks')}, d['e2e'], d['cpu_baseline'], d['roofline'], d['roofline_range_count'])
print(d['stages_ms'])
PY
for tool in memcheck racecheck synccheck; do echo "== $tool"; timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_dbscan.py tests/test_gpu_scoring.py tests/test_gpu_edge.py tests/test_gpu_rounds.py -m gpu -q -x -k "battery_labels or counts_battery or range_counts_random or summary or cut_rounds or counting_sort or tiny or rounds_match" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|hazard|Invalid|error" | head -8; done
