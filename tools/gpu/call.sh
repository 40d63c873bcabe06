set -x
timeout 600 python -m pytest tests/test_gpu_scheduler.py -m gpu -q -x > gpurun_out/s34_tests.log 2>&1; tail -3 gpurun_out/s34_tests.log
timeout 300 python bench.py --steps 100 --no-cpu-baseline > gpurun_out/s34_n1.json 2> gpurun_out/s34_n1.err
TR="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/s34_n2.json 2> gpurun_out/s34_n2.err; tail -c 300 gpurun_out/s34_n2.err
python - <<PY
import json
for f in ('s34_n1','s34_n2'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    print(d['n_gpus'], '%.1fM'%(d['value']/1e6), d['ms_per_step'], 'e2e %.1fM'%(d['e2e']['value']/1e6), d['e2e']['ms_per_step'], d['e2e']['host_ms_median'])
PY
