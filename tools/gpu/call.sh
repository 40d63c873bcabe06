set -x
python -m pytest tests -m gpu -x -q > gpurun_out/s3_tests.log 2>&1; tail -3 gpurun_out/s3_tests.log
python bench.py --model teecnet --precision f16 --steps 20 --no-cpu-baseline > gpurun_out/s3_teec_f16.json 2> gpurun_out/s3_teec_f16.err; tail -c 600 gpurun_out/s3_teec_f16.err
python bench.py --model teecnet --precision tf32 --steps 20 --no-cpu-baseline > gpurun_out/s3_teec_tf32.json 2> gpurun_out/s3_teec_tf32.err
python tools/bench_train.py --mesh-n 28 --precision tf32 --steps 5 --profile > gpurun_out/s3_train28.json 2> gpurun_out/s3_train28.err; cat gpurun_out/s3_train28.json
python tools/bench_train.py --mesh-n 28 --precision tf32 --steps 5 --model teecnet --profile > gpurun_out/s3_train28_teec.json 2> gpurun_out/s3_train28_teec.err; cat gpurun_out/s3_train28_teec.json
python tools/bench_train.py --mesh-n 44 --levels 9 --precision tf32 --steps 5 --profile > gpurun_out/s3_train44.json 2> gpurun_out/s3_train44.err; cat gpurun_out/s3_train44.json
