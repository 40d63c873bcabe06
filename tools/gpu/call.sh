set -x
python -m pytest tests/test_gpu_forward.py tests/test_gpu_backward.py tests/test_gpu_scheduler.py -m gpu -q -s > gpurun_out/s5_tests.log 2>&1; tail -3 gpurun_out/s5_tests.log; grep "rel-L2" gpurun_out/s5_tests.log
for prec in f16 tf32; do
python bench.py --model teecnet --precision $prec --steps 20 --no-cpu-baseline > gpurun_out/s5_teec_$prec.json 2> gpurun_out/s5_teec_$prec.err; tail -c 300 gpurun_out/s5_teec_$prec.err
python - <<PY
import json
d=json.loads(open('gpurun_out/s5_teec_$prec.json').read().strip().splitlines()[-1])
print('%.1fM'%(d['value']/1e6), d['ms_per_step'], {k:round(v['ms_per_launch'],3) for k,v in d['kernels'].items()})
PY
done
