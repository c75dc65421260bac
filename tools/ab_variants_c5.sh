#!/bin/bash
# A/B over gpurun_variants/*.so on config 5 (3M Gaussians, 1024x2048, forward only): stage timers per frame
mkdir -p gpurun_out
OUT=gpurun_out/${1:-ab_variants_c5}.txt
for lib in "" gpurun_variants/*.so ""; do
  if [ -z "$lib" ]; then name=default; unset S360_LIB; else name=$lib; export S360_LIB=$PWD/$lib; fi
  timeout -s KILL 300 python bench.py --config 5 --steps 20 --warmup 3 --no-e2e 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); s=d['roofline']['stages_ms']
print('$name', 'ms/step %.4f'%d['ms_per_step'], ' '.join('%s=%.4f'%(k[:12],v) for k,v in s.items()))"
done 2>&1 | tee $OUT
