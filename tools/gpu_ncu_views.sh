#!/bin/bash
# ncu --set full of the batched-path kernels: usage gpu_ncu_views.sh <kernel regex> <count> <outname> [skip]
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$1" -s ${4:-4} -c $2 -o gpurun_out/$3 -f python tools/views_bench.py 256 --no-separate > gpurun_out/$3.log 2>&1
tail -2 gpurun_out/$3.log
