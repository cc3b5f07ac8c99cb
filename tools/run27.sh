set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for s in 6 8; do
CLOOPS_STREAMS=$s timeout 200 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_c4_streams$s.json 2> gpurun_out/err_s$s.txt
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_c4_streams$s.json').read().strip().splitlines()[-1])
print('STREAMS=$s', d['ms_per_step'], d['e2e']['ms_per_step'])
PY
done
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_c4_streams4b.json 2> gpurun_out/err_s4.txt
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_c4_streams4b.json').read().strip().splitlines()[-1])
print('STREAMS=4', d['ms_per_step'], d['e2e']['ms_per_step'])
PY
