#!/bin/bash
# Graphed decoder step: new tests, the whole GPU suite, config 4 with and without the graph, headline bench.
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_graph.py -m gpu -x -q 2>&1 | tail -n 25 > gpurun_out/r02k_graph_tests.log
timeout -s KILL 900 python -m pytest tests -m gpu -q 2>&1 | tail -n 15 > gpurun_out/r02k_pytest.log
timeout -s KILL 600 python bench.py --config 4 --steps 30 --warmup 5 > gpurun_out/r02k_c4_graph.json 2> gpurun_out/r02k_c4_graph.err
timeout -s KILL 600 python bench.py --config 4 --steps 30 --warmup 5 --no-graph > gpurun_out/r02k_c4_eager.json 2> gpurun_out/r02k_c4_eager.err
timeout -s KILL 600 python bench.py --steps 100 --warmup 5 --no-cube6 > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err
cat gpurun_out/r02k_graph_tests.log; tail -n 6 gpurun_out/r02k_pytest.log
for f in gpurun_out/r02k_c4_graph.json gpurun_out/r02k_c4_eager.json gpurun_out/r02k_bench.json; do echo $f; python -c "
import json
try:
  d=json.loads(open('$f').read().strip().splitlines()[-1])
  print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches')}, d['config'].get('graph'))
except Exception as e: print('ERR', e)
"; done
tail -n 5 gpurun_out/r02k_c4_graph.err gpurun_out/r02k_c4_eager.err gpurun_out/r02k_bench.err
