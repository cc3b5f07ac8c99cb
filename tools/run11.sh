set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r02_launches_step.csv python tools/ncu_step.py > gpurun_out/ncu_step.log 2>&1
tail -2 gpurun_out/ncu_step.log
python tools/launch_shares2.py gpurun_out/r02_launches_step.csv gpurun_out/r02_step_kernel_shares.csv; head -25 gpurun_out/r02_step_kernel_shares.csv
for c in 2 3 5; do
timeout 900 python bench.py --config $c --steps 3 --warmup 3 > gpurun_out/r02_bench_c$c.json 2> gpurun_out/r02_bench_c$c.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_c$c.json').read().strip().splitlines()[-1])
print($c, {k:d[k] for k in ('value','ms_per_step','gpu_launches','result')}, d['e2e']['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d.get('cpu_baseline',{}).get('value'))
PY
tail -2 gpurun_out/r02_bench_c$c.err
done
