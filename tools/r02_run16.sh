#!/bin/bash
# Fused batched K1: views / binning / graph tests first, then the whole suite, config 4 and the six-face side bench.
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_views.py tests/test_gpu_graph.py tests/test_gpu_baseline_configs.py -m gpu -x -q 2>&1 | tail -n 25 > gpurun_out/r02o_views_tests.log
timeout -s KILL 900 python -m pytest tests -m gpu -q 2>&1 | tail -n 12 > gpurun_out/r02o_pytest.log
timeout -s KILL 600 python bench.py --config 4 --steps 50 --warmup 5 > gpurun_out/r02o_c4.json 2> gpurun_out/r02o_c4.err
timeout -s KILL 300 python tools/views_bench.py 256 > gpurun_out/r02o_views.json 2>&1
timeout -s KILL 300 python tools/profile_config4.py > gpurun_out/r02o_profile_config4.log 2>&1
tail -n 8 gpurun_out/r02o_views_tests.log; tail -n 5 gpurun_out/r02o_pytest.log
python -c "
import json
d=json.loads(open('gpurun_out/r02o_c4.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches')}, {k:round(v,4) for k,v in d['roofline']['stages_ms'].items()})
"
tail -n 3 gpurun_out/r02o_c4.err; tail -c 700 gpurun_out/r02o_views.json
sed -n '/device time per step/,$p' gpurun_out/r02o_profile_config4.log | head -12
