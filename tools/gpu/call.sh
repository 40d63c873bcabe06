set -x
timeout 600 python -m pytest tests/test_gpu_interp.py -m gpu -q -x > gpurun_out/s17_tests.log 2>&1; tail -15 gpurun_out/s17_tests.log
TR="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29515 tools/bench_train.py --mesh-n 28 --precision tf32 --steps 5 > gpurun_out/s17_train_n2.json 2> gpurun_out/s17_train_n2.err; cat gpurun_out/s17_train_n2.json; tail -c 300 gpurun_out/s17_train_n2.err
