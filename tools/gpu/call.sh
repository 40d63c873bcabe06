set -x
N=8
TR="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29511 bench.py --gpus $N --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/s35_n8.json 2> gpurun_out/s35_n8.err; tail -c 300 gpurun_out/s35_n8.err
python - <<PY
import json
d=json.loads(open('gpurun_out/s35_n8.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['config']['cells'], '%.1fM'%(d['value']/1e6), d['ms_per_step'], 'e2e %.1fM'%(d['e2e']['value']/1e6), d['e2e']['ms_per_step'], d['e2e']['host_ms_median'])
PY
