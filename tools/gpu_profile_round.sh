#!/bin/bash
# Round profile: bench line, ncu launch list of the same command, ncu --set full of every library kernel.
R=${1:-r01}
mkdir -p gpurun_out
[ "$2" == "ncuonly" ] || python bench.py --steps 200 --warmup 10 > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err
[ "$2" == "ncuonly" ] || python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/${R}_bench_reference.json 2>> gpurun_out/${R}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'s360|render_|preprocess_|rs_|scan_|emit_|tile_|mark_' --launch-skip 52 --launch-count 80 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu --no-cube6 > gpurun_out/${R}_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'s360|render_|preprocess_|rs_|scan_|emit_|tile_' --launch-skip 52 --launch-count 14 -o gpurun_out/${R}_full -f python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-cube6 > gpurun_out/${R}_ncu_full.log 2>&1
tail -2 gpurun_out/${R}_bench.err; cut -c1-400 gpurun_out/${R}_bench.json; cut -c1-300 gpurun_out/${R}_bench_reference.json
