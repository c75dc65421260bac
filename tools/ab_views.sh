#!/bin/bash
# A/B of variant libraries (gpurun_variants/*.so) on the batched six-face pass
mkdir -p gpurun_out
for lib in "" gpurun_variants/*.so; do
  [ "$lib" == "gpurun_variants/*.so" ] && continue
  if [ -z "$lib" ]; then unset S360_LIB; extra=""; else export S360_LIB=$PWD/$lib; extra="--no-separate"; fi
  timeout -s KILL 300 python tools/views_bench.py 256 $extra 2>&1 | tail -1 | tee -a gpurun_out/ab_views.log
done
