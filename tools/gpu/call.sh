set -x
timeout 600 python -m pytest tests/test_gpu_backward.py -m gpu -q -x > gpurun_out/s24_tests.log 2>&1; tail -3 gpurun_out/s24_tests.log
python tools/bench_train.py --mesh-n 28 --precision tf32 --steps 5 > gpurun_out/s24_train28.json 2> gpurun_out/s24_train28.err; cat gpurun_out/s24_train28.json; tail -c 300 gpurun_out/s24_train28.err
python tools/bench_train.py --mesh-n 28 --precision tf32 --steps 5 --model teecnet > gpurun_out/s24_train28_teec.json 2> gpurun_out/s24_train28_teec.err; cat gpurun_out/s24_train28_teec.json
