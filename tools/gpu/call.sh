set -x
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/s20_tests.log 2>&1; tail -4 gpurun_out/s20_tests.log
timeout 300 python bench.py --steps 100 --no-cpu-baseline > gpurun_out/s20_bench.json 2> gpurun_out/s20_bench.err; tail -c 300 gpurun_out/s20_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/s20_bench.json').read().strip().splitlines()[-1])
print('%.1fM'%(d['value']/1e6), d['ms_per_step'], 'e2e %.1fM'%(d['e2e']['value']/1e6), d['e2e']['ms_per_step'], d['e2e']['host_ms_median'])
PY
