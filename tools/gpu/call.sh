set -x
timeout 900 python -m pytest tests -m gpu -q -x -s > gpurun_out/s22_tests.log 2>&1; tail -3 gpurun_out/s22_tests.log; grep "teecnet 500k" gpurun_out/s22_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 400 python bench.py > gpurun_out/s22_bench.json 2> gpurun_out/s22_bench.err; tail -c 300 gpurun_out/s22_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/s22_bench.json').read().strip().splitlines()[-1])
print('%.1fM'%(d['value']/1e6), d['ms_per_step'], 'e2e %.1fM'%(d['e2e']['value']/1e6), d['roofline'], d['cpu_baseline'], d['clocks'])
PY
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s22_ref.json 2> gpurun_out/s22_ref.err; cut -c 1-300 gpurun_out/s22_ref.json
