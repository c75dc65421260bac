#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_graph.py tests/test_gpu_api.py tests/test_gpu_views.py -m gpu -q 2>&1 | tail -n 25 > gpurun_out/r02l_tests.log
timeout -s KILL 600 python bench.py --config 4 --steps 30 --warmup 5 > gpurun_out/r02l_c4_graph.json 2> gpurun_out/r02l_c4_graph.err
timeout -s KILL 600 python bench.py --config 4 --steps 30 --warmup 5 --no-graph > gpurun_out/r02l_c4_eager.json 2> gpurun_out/r02l_c4_eager.err
timeout -s KILL 600 python tools/profile_config4.py > gpurun_out/r02l_profile_config4.log 2>&1
tail -n 8 gpurun_out/r02l_tests.log
for f in gpurun_out/r02l_c4_graph.json gpurun_out/r02l_c4_eager.json; do echo $f; python -c "
import json
try:
  d=json.loads(open('$f').read().strip().splitlines()[-1])
  print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches')}, d['config'].get('graph'))
except Exception as e: print('ERR', e)
"; done
tail -n 3 gpurun_out/r02l_c4_graph.err gpurun_out/r02l_c4_eager.err
sed -n '/device time per step/,$p' gpurun_out/r02l_profile_config4.log | head -50
