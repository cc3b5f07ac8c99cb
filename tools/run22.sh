set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --config 5 > gpurun_out/r02_bench_c5.json 2> gpurun_out/r02_bench_c5.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_c5.json').read().strip().splitlines()[-1])
print(5, {k:d[k] for k in ('value','ms_per_step','gpu_launches','result')}, d['e2e']['ms_per_step'], d['roofline']['frac'], d.get('cpu_baseline'))
PY
tail -3 gpurun_out/r02_bench_c5.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/r02_launches_step.csv python tools/ncu_step.py 4 200000000 3 > gpurun_out/ncu_step.log 2>&1
tail -2 gpurun_out/ncu_step.log
python tools/launch_shares2.py gpurun_out/r02_launches_step.csv gpurun_out/r02_step_kernel_shares.csv; head -30 gpurun_out/r02_step_kernel_shares.csv
