set -x
timeout 600 python -m pytest tests/test_gpu_forward.py tests/test_gpu_scheduler.py tests/test_gpu_backward.py -m gpu -q -x -s > gpurun_out/s21_tests.log 2>&1; tail -4 gpurun_out/s21_tests.log; grep "teecnet" gpurun_out/s21_tests.log
timeout 300 python bench.py --model teecnet --steps 30 --no-cpu-baseline > gpurun_out/s21_teec.json 2> gpurun_out/s21_teec.err; tail -c 300 gpurun_out/s21_teec.err
python - <<PY
import json
d=json.loads(open('gpurun_out/s21_teec.json').read().strip().splitlines()[-1])
print('%.1fM'%(d['value']/1e6), d['ms_per_step'], 'e2e %.1fM'%(d['e2e']['value']/1e6), {k:round(v['ms_per_launch'],4) for k,v in d['kernels'].items()})
PY
