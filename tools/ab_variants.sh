#!/bin/bash
# A/B over every variant library in gpurun_variants/ (tools/build_variants.py): default first and last; stage timers of config 3.
mkdir -p gpurun_out
OUT=gpurun_out/${1:-ab_variants}.txt
for lib in "" gpurun_variants/*.so ""; do
  if [ -z "$lib" ]; then name=default; unset S360_LIB; else name=$lib; export S360_LIB=$PWD/$lib; fi
  timeout -s KILL 300 python bench.py --steps 100 --warmup 5 --no-e2e --no-cpu --no-cube6 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); s=d['roofline']['stages_ms']
print('$name', 'ms/step %.4f'%d['ms_per_step'], ' '.join('%s=%.4f'%(k[:12],v) for k,v in s.items()))"
done 2>&1 | tee $OUT
