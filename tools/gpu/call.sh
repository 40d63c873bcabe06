set -x
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x -s > gpurun_out/s16_tests.log 2>&1; tail -15 gpurun_out/s16_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
