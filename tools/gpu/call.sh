set -x
CMD="python tools/bench_train.py --mesh-n 28 --precision tf32 --steps 2 --warmup 1 --model teecnet"
$CMD > gpurun_out/r01c_train_teec_plain.json 2> gpurun_out/r01c_train_teec_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r01c_train_teec_launches.csv $CMD > gpurun_out/r01c_train_teec_ncu.log 2>&1
tail -1 gpurun_out/r01c_train_teec_ncu.log
