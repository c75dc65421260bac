#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'s360|render_|preprocess_|rs_|scan_|emit_|tile_ranges' --launch-skip 120 --launch-count 40 -o gpurun_out/prof_all -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-cube6 > gpurun_out/prof_all.log 2>&1
tail -2 gpurun_out/prof_all.log
