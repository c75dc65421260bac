#!/bin/bash
# Round-2 GPU session 1: full GPU test suite (all failures listed), work counters, A/B of prepared variants.
mkdir -p gpurun_out
timeout -s KILL 2000 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02a_pytest.log
S360_LIB=$PWD/gpurun_variants/lib_counters.so timeout -s KILL 300 python tools/counters.py > gpurun_out/r02a_counters.json 2> gpurun_out/r02a_counters.err
line() { python -c "
import sys,json
d=json.loads(sys.stdin.read()); s=d['roofline']['stages_ms']
print('$1', 'ms/step %.3f'%d['ms_per_step'], ' '.join('%s=%.3f'%(k[:12],v) for k,v in s.items()))"; }
ab() {
  S360_FORCE_RADIX_BINNING=1 timeout -s KILL 300 python bench.py --steps 60 --warmup 5 --no-e2e --no-cpu --no-cube6 2>/dev/null | line radix_binning
  for lib in "" gpurun_variants/lib_*.so; do
    case "$lib" in *counters*) continue;; esac
    if [ -z "$lib" ]; then name=default; unset S360_LIB; else name=$lib; export S360_LIB=$PWD/$lib; fi
    timeout -s KILL 300 python bench.py --steps 60 --warmup 5 --no-e2e --no-cpu --no-cube6 2>/dev/null | line $name
  done
  unset S360_LIB
}
ab > gpurun_out/r02a_ab.log 2>&1
S360_LIB=$PWD/gpurun_variants/lib_qprime.so timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_views.py tests/test_gpu_baseline_configs.py -m gpu -q 2>&1 | tail -15 > gpurun_out/r02a_pytest_qprime.log
timeout -s KILL 600 python bench.py --steps 100 --warmup 10 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
cat gpurun_out/r02a_pytest.log; cat gpurun_out/r02a_counters.json; cat gpurun_out/r02a_ab.log; cat gpurun_out/r02a_pytest_qprime.log; tail -3 gpurun_out/r02a_bench.err; cut -c1-1500 gpurun_out/r02a_bench.json
