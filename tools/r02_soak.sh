#!/bin/bash
# Soak: long runs of every bench configuration (graph replay and eager), looking for hangs / overflow flags / drift.
mkdir -p gpurun_out
: > gpurun_out/r02s_soak.txt
run() { local tag=$1; shift; local t0=$(date +%s.%N); timeout -s KILL 240 python bench.py "$@" > /tmp/soak.json 2> /tmp/soak.err; local rc=$?; local t1=$(date +%s.%N)
  python - "$tag" $rc $t0 $t1 <<'PY' >> gpurun_out/r02s_soak.txt
import json, sys
tag, rc, t0, t1 = sys.argv[1], int(sys.argv[2]), float(sys.argv[3]), float(sys.argv[4])
try:
    d = json.loads(open('/tmp/soak.json').read().strip().splitlines()[-1])
    print(tag, 'rc', rc, 'wall %.1fs' % (t1 - t0), 'steps', d['steps'], 'ms/step %.4f' % d['ms_per_step'], 'clocks', (d.get('clocks') or {}).get('sm_mhz'), (d.get('clocks') or {}).get('reasons'))
except Exception as e:
    print(tag, 'rc', rc, 'wall %.1fs' % (t1 - t0), 'ERR', e, open('/tmp/soak.err').read()[-400:])
PY
}
for rep in 1 2; do
  run c3_graph_$rep --steps 5000 --warmup 5 --no-e2e --no-cpu --no-cube6
  run c3_eager_$rep --steps 3000 --warmup 5 --no-e2e --no-cpu --no-cube6 --no-graph
  run c3_exact_$rep --steps 1500 --warmup 5 --no-e2e --no-cpu --no-cube6 --no-graph --exact-counts
  run c4_graph_$rep --config 4 --steps 3000 --warmup 5
  run c4_eager_$rep --config 4 --steps 1500 --warmup 5 --no-graph
  run c5_$rep --config 5 --steps 300 --warmup 3 --no-e2e
done
cat gpurun_out/r02s_soak.txt
