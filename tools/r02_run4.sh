#!/bin/bash
# Round-2 GPU session 4: full GPU suite, A/B, graph vs autograd step, ncu launch list + full capture, bench.
mkdir -p gpurun_out
timeout -s KILL 2400 python -m pytest tests -m gpu -q -rs 2>&1 > gpurun_out/r02d_pytest_full.log; tail -30 gpurun_out/r02d_pytest_full.log > gpurun_out/r02d_pytest.log
line() { python -c "
import sys,json
d=json.loads(sys.stdin.read()); s=d['roofline']['stages_ms']
print('$1', 'ms/step %.3f'%d['ms_per_step'], 'sum %.3f'%sum(s.values()), ' '.join('%s=%.3f'%(k[:12],v) for k,v in s.items()))"; }
{ for lib in "" gpurun_variants/lib_unstaged.so; do
    if [ -z "$lib" ]; then name=default; unset S360_LIB; else name=$lib; export S360_LIB=$PWD/$lib; fi
    timeout -s KILL 300 python bench.py --steps 60 --warmup 5 --no-e2e --no-cpu --no-cube6 --no-graph 2>/dev/null | line $name
  done
  unset S360_LIB
  timeout -s KILL 300 python bench.py --steps 60 --warmup 5 --no-e2e --no-cpu --no-cube6 2>/dev/null | line graph_step
} > gpurun_out/r02d_ab.log 2>&1
timeout -s KILL 900 python bench.py --steps 100 --warmup 10 > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err
cat gpurun_out/r02d_pytest.log; cat gpurun_out/r02d_ab.log; tail -3 gpurun_out/r02d_bench.err; head -c 600 gpurun_out/r02d_bench.json; python -c "
import json; d=json.load(open('gpurun_out/r02d_bench.json')); print(json.dumps(d['parity'])[:1500]); print(d['e2e'])"
