#!/bin/bash
# A/B: resident CTAs per SM of the render kernels (launch bounds 8 -> 62 / 64 registers, no spills)
mkdir -p gpurun_out
for lib in "" gpurun_variants/lib_bwd8.so gpurun_variants/lib_fwd7.so gpurun_variants/lib_fwd8.so gpurun_variants/lib_both8.so ""; do
  if [ -z "$lib" ]; then name=default; unset S360_LIB; else name=$lib; export S360_LIB=$PWD/$lib; fi
  timeout -s KILL 300 python bench.py --steps 100 --warmup 5 --no-e2e --no-cpu --no-cube6 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); s=d['roofline']['stages_ms']
print('$name', 'ms/step %.4f'%d['ms_per_step'], ' '.join('%s=%.4f'%(k[:12],v) for k,v in s.items()), 'parity', (d.get('parity') or {}).get('ok'))"
done 2>&1 | tee gpurun_out/r02p_ab_occupancy.txt
