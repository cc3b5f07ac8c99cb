set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
CLOOPS_TRACE=1 timeout 600 python tools/pass_probe.py 0 > gpurun_out/r02_probe_chr1.log 2>&1
grep -v "^\[cloops\]" gpurun_out/r02_probe_chr1.log | tail -20
grep "^\[cloops\]" gpurun_out/r02_probe_chr1.log | sort | uniq -c | sort -rn | head -12
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_pipe.py tests/test_gpu_scoring.py -x -q 2>&1 | tail -8 > gpurun_out/r02_pytest2.log
tail -8 gpurun_out/r02_pytest2.log
timeout 900 python bench.py --config 4 --steps 3 --warmup 3 > gpurun_out/r02_bench_c4b.json 2> gpurun_out/r02_bench_c4b.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_c4b.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','result')}, d['e2e'], d['cpu_baseline'])
print(d['stages_ms'])
PY
tail -3 gpurun_out/r02_bench_c4b.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches_chr21.csv python tools/pass_probe.py 20 > gpurun_out/r02_probe_chr21_ncu.log 2>&1
tail -4 gpurun_out/r02_probe_chr21_ncu.log
