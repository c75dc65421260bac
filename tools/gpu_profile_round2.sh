#!/bin/bash
# Round profile, second half of round 1: everything of gpu_profile_round.sh plus the batched six-face pass.
R=${1:-r01}
mkdir -p gpurun_out
bash tools/gpu_profile_round.sh $R
python tools/views_bench.py 256 > gpurun_out/${R}_views_bench.json 2> gpurun_out/${R}_views_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'s360|render_|preprocess_|multi_|rs_|scan_|emit_|tile_|zero_acc|cube2' --launch-skip 24 --launch-count 48 --csv --log-file gpurun_out/${R}_views_launches.csv python tools/views_bench.py 256 --profile > gpurun_out/${R}_views_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'multi_|render_|zero_acc' --launch-skip 12 --launch-count 12 -o gpurun_out/${R}_views_full -f python tools/views_bench.py 256 --profile > gpurun_out/${R}_views_ncu_full.log 2>&1
timeout 900 python tests/tools/run_configs.py > gpurun_out/${R}_configs.log 2>&1
cp gpurun_out/configs.json gpurun_out/${R}_configs.json
tail -1 gpurun_out/${R}_views_bench.json | cut -c1-700
