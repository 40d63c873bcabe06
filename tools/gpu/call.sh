set -x
python -m pytest tests/test_gpu_scheduler.py -m gpu -q -x > gpurun_out/s9_tests.log 2>&1; tail -3 gpurun_out/s9_tests.log
TR="timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29513 tools/check_alds_multi.py --mesh-n 28 --clusters 4 --model teecnet > gpurun_out/s9_alds4.json 2> gpurun_out/s9_alds4.err; cat gpurun_out/s9_alds4.json; tail -c 300 gpurun_out/s9_alds4.err
$TR --master-port 29514 tools/check_alds_multi.py --mesh-n 28 --clusters 4 --model neuralop > gpurun_out/s9_alds4k.json 2> gpurun_out/s9_alds4k.err; cat gpurun_out/s9_alds4k.json; tail -c 300 gpurun_out/s9_alds4k.err
