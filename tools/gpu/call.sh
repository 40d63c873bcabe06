set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/s7_n2.json 2> gpurun_out/s7_n2.err; tail -c 400 gpurun_out/s7_n2.err
python - <<PY
import json
d=json.loads(open('gpurun_out/s7_n2.json').read().strip().splitlines()[-1])
print('%.1fM'%(d['value']/1e6), d['ms_per_step'], 'e2e %.1fM'%(d['e2e']['value']/1e6))
PY
$TR --master-port 29512 tools/check_alds_multi.py --mesh-n 28 --clusters 4 > gpurun_out/s7_alds4.json 2> gpurun_out/s7_alds4.err; cat gpurun_out/s7_alds4.json; tail -c 600 gpurun_out/s7_alds4.err
$TR --master-port 29513 tools/check_alds_multi.py --mesh-n 28 --clusters 1 --model teecnet > gpurun_out/s7_alds1.json 2> gpurun_out/s7_alds1.err; cat gpurun_out/s7_alds1.json; tail -c 600 gpurun_out/s7_alds1.err
