#!/bin/bash
# rs10 (10 keys per thread in the onesweep passes) against the default on configs 4 / 5 and through the exact-list tests
mkdir -p gpurun_out
S360_LIB=$PWD/gpurun_variants/lib_rs10.so timeout -s KILL 900 python -m pytest tests/test_gpu_stages.py tests/test_gpu_binning_paths.py tests/test_gpu_views.py tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -n 6
for lib in "" gpurun_variants/lib_rs10.so ""; do
  if [ -z "$lib" ]; then name=default; unset S360_LIB; else name=$lib; export S360_LIB=$PWD/$lib; fi
  for cfg in 4 5; do
  timeout -s KILL 300 python bench.py --config $cfg --steps 30 --warmup 5 --no-e2e 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); s=d['roofline']['stages_ms']
print('$name', 'config $cfg', 'ms/step %.4f'%d['ms_per_step'], ' '.join('%s=%.4f'%(k[:12],v) for k,v in s.items()))"
  done
done 2>&1 | tee gpurun_out/r02q_ab_rs10_c45.txt
