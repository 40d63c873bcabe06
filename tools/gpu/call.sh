set -x
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/r01c_plain.json 2> gpurun_out/r01c_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r01c_launches.csv $CMD > gpurun_out/r01c_ncu1.log 2>&1
$CMD > gpurun_out/r01c_plain2.json 2> gpurun_out/r01c_plain2.err &&
ncu --set full --clock-control none --import-source on -k regex:layer_fused -s 10 -c 2 -o gpurun_out/r01c_top -f $CMD > gpurun_out/r01c_ncu2.log 2>&1
ls -la gpurun_out/ | tail -8
