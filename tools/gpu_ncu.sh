#!/bin/bash
# ncu full capture of selected kernels: usage gpu_ncu.sh <regex> <count> <outname>
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$1" -s ${4:-8} -c $2 -o gpurun_out/$3 -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-cube6 > gpurun_out/$3.log 2>&1
tail -2 gpurun_out/$3.log
