set -x
timeout 600 python -m pytest tests/test_gpu_backward.py -m gpu -q -x > gpurun_out/s32_tests.log 2>&1; tail -3 gpurun_out/s32_tests.log
for m in neuralop teecnet; do
python tools/bench_train.py --mesh-n 28 --precision tf32 --steps 5 --model $m > gpurun_out/s32_train28_$m.json 2> gpurun_out/s32_train28.err; python -c "
import json;d=json.loads(open('gpurun_out/s32_train28_$m.json').read().strip().splitlines()[-1]);print('$m', d['ms_per_step'], d['value'], d['loss'])"
done
