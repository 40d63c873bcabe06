set -x
timeout 600 python -m pytest tests/test_gpu_backward.py -m gpu -q -x > gpurun_out/s25_tests.log 2>&1; tail -3 gpurun_out/s25_tests.log
for c in 4 6 2; do
FESR_WGRAD_CTAS=$c python tools/bench_train.py --mesh-n 28 --precision tf32 --steps 5 > gpurun_out/s25_train28_$c.json 2> gpurun_out/s25_train28.err; python -c "
import json;d=json.loads(open('gpurun_out/s25_train28_$c.json').read().strip().splitlines()[-1]);print($c, d['ms_per_step'])"
done
