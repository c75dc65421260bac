#!/bin/bash
# A/B: run the bench stage timings for each variant library in gpurun_variants/
timeout -s KILL 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for lib in "" gpurun_variants/*.so; do
  [ "$lib" == "gpurun_variants/*.so" ] && continue
  if [ -z "$lib" ]; then name=default; unset S360_LIB; else name=$lib; export S360_LIB=$PWD/$lib; fi
  timeout -s KILL 300 python bench.py --steps 60 --warmup 5 --no-e2e --no-cpu --no-cube6 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); s=d['roofline']['stages_ms']
print('$name', 'ms/step %.3f'%d['ms_per_step'], ' '.join('%s=%.3f'%(k[:12],v) for k,v in s.items()))"
done
