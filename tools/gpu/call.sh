set -x
timeout 600 python -m pytest tests/test_gpu_backward.py -m gpu -q -x > gpurun_out/s27_tests.log 2>&1; tail -5 gpurun_out/s27_tests.log
python tools/bench_train.py --mesh-n 28 --precision tf32 --steps 5 > gpurun_out/s27_train28.json 2> gpurun_out/s27_train28.err; python -c "
import json;d=json.loads(open('gpurun_out/s27_train28.json').read().strip().splitlines()[-1]);print(d['ms_per_step'], d['value'], d['loss'])"
FESR_EDGE_MLP_BWD_GENERIC=1 python tools/bench_train.py --mesh-n 28 --precision tf32 --steps 5 > gpurun_out/s27_train28_g.json 2> gpurun_out/s27_train28_g.err; python -c "
import json;d=json.loads(open('gpurun_out/s27_train28_g.json').read().strip().splitlines()[-1]);print(d['ms_per_step'], d['value'], d['loss'])"
