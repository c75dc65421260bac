#!/bin/bash
# One GPU session: tests, smoke, bench, ncu launch list, ncu full capture of the render kernels.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
python bench.py --steps 100 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err
if [ "$1" != "quick" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed/" -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-cube6 > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "timed/" -k regex:'render_|preprocess_' -c 4 -o gpurun_out/prof_render -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-cube6 > gpurun_out/ncu_full.log 2>&1
fi
cat gpurun_out/pytest_gpu.log gpurun_out/smoke.log; tail -3 gpurun_out/bench.err; cut -c1-3000 gpurun_out/bench.json
