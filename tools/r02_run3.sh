#!/bin/bash
# Round-2 GPU session 3: full GPU suite, A/B of record staging / atan2 / match variants, graph vs autograd step, bench.
mkdir -p gpurun_out
timeout -s KILL 2400 python -m pytest tests -m gpu -q -rs 2>&1 > gpurun_out/r02c_pytest_full.log; tail -30 gpurun_out/r02c_pytest_full.log > gpurun_out/r02c_pytest.log
line() { python -c "
import sys,json
d=json.loads(sys.stdin.read()); s=d['roofline']['stages_ms']
print('$1', 'ms/step %.3f'%d['ms_per_step'], 'sum %.3f'%sum(s.values()), ' '.join('%s=%.3f'%(k[:12],v) for k,v in s.items()))"; }
{ for lib in "" gpurun_variants/lib_unstaged.so gpurun_variants/lib_atan2f.so gpurun_variants/lib_mbballot.so; do
    if [ -z "$lib" ]; then name=default; unset S360_LIB; else name=$lib; export S360_LIB=$PWD/$lib; fi
    timeout -s KILL 300 python bench.py --steps 60 --warmup 5 --no-e2e --no-cpu --no-cube6 --no-graph 2>/dev/null | line $name
  done
  unset S360_LIB
  S360_FORCE_RADIX_BINNING=1 timeout -s KILL 300 python bench.py --steps 60 --warmup 5 --no-e2e --no-cpu --no-cube6 --no-graph 2>/dev/null | line radix_binning
  timeout -s KILL 300 python bench.py --steps 60 --warmup 5 --no-e2e --no-cpu --no-cube6 2>/dev/null | line graph_step
} > gpurun_out/r02c_ab.log 2>&1
S360_LIB=$PWD/gpurun_variants/lib_atan2f.so timeout -s KILL 600 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -q -k "config1 or config3_1m_pixel_aligned_512" 2>&1 | grep -E "rel-L2|passed|failed" > gpurun_out/r02c_pytest_atan2f.log
timeout -s KILL 900 python bench.py --steps 100 --warmup 10 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err
cat gpurun_out/r02c_pytest.log; cat gpurun_out/r02c_ab.log; cat gpurun_out/r02c_pytest_atan2f.log; tail -3 gpurun_out/r02c_bench.err; head -c 1200 gpurun_out/r02c_bench.json
