set -x
timeout 300 python -m pytest tests/test_gpu_forward.py -m gpu -q -x -k "fused or prepared or 50k or deterministic" > gpurun_out/s11_tests.log 2>&1; tail -3 gpurun_out/s11_tests.log
for pdl in 1 0; do
FESR_PDL=$pdl timeout 300 python bench.py --steps 100 --no-cpu-baseline > gpurun_out/s11_bench_pdl$pdl.json 2> gpurun_out/s11_bench_pdl$pdl.err; tail -c 300 gpurun_out/s11_bench_pdl$pdl.err
python - <<PY
import json
d=json.loads(open('gpurun_out/s11_bench_pdl$pdl.json').read().strip().splitlines()[-1])
print('pdl$pdl %.1fM'%(d['value']/1e6), d['ms_per_step'], d['ms_per_step_instrumented'], 'e2e %.1fM'%(d['e2e']['value']/1e6), {k:round(v['ms_per_launch'],4) for k,v in d['kernels'].items()})
PY
done
