set -x
timeout 600 python -m pytest tests/test_gpu_scheduler.py -m gpu -q -x > gpurun_out/s15_tests.log 2>&1; tail -5 gpurun_out/s15_tests.log
TR="timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/s15_n2.json 2> gpurun_out/s15_n2.err; tail -c 400 gpurun_out/s15_n2.err
python - <<PY
import json
d=json.loads(open('gpurun_out/s15_n2.json').read().strip().splitlines()[-1])
print('%.1fM'%(d['value']/1e6), d['ms_per_step'], 'e2e %.1fM'%(d['e2e']['value']/1e6), d['e2e'])
PY
$TR --master-port 29513 tools/check_alds_multi.py --mesh-n 28 --clusters 1 --model neuralop > gpurun_out/s15_alds1.json 2> gpurun_out/s15_alds1.err; cat gpurun_out/s15_alds1.json; tail -c 300 gpurun_out/s15_alds1.err
