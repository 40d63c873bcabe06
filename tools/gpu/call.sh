set -x
N=4
TR="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/s30_n4.json 2> gpurun_out/s30_n4.err; tail -c 300 gpurun_out/s30_n4.err
$TR --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/s30_ref_n4.json 2> gpurun_out/s30_ref_n4.err; cut -c 1-200 gpurun_out/s30_ref_n4.json
$TR --master-port 29513 tools/check_alds_multi.py --mesh-n 28 --clusters 4 --model teecnet > gpurun_out/s30_alds4.json 2> gpurun_out/s30_alds4.err; cat gpurun_out/s30_alds4.json; tail -c 200 gpurun_out/s30_alds4.err
python - <<PY
import json
d=json.loads(open('gpurun_out/s30_n4.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['config']['cells'], '%.1fM'%(d['value']/1e6), d['ms_per_step'], 'e2e %.1fM'%(d['e2e']['value']/1e6), d['e2e']['ms_per_step'], d['cpu_baseline'])
PY
