#!/bin/bash
# ncu evidence per the profiling recipe: launch list (time shares) + one full capture of the dominant kernel.
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 140 -c 4 -o gpurun_out/conv_full $CMD > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out
