#!/bin/bash
# Round-2 GPU session 6: GPU suite, A/B of async staging (render) and of the batched K8+K9 launch bound, configs 4/5 at N=1.
mkdir -p gpurun_out
timeout -s KILL 2400 python -m pytest tests -m gpu -q -rs -x 2>&1 | tail -8 > gpurun_out/r02f_pytest.log
line() { python -c "
import sys,json
d=json.loads(sys.stdin.read()); s=d['roofline']['stages_ms']
print('$1', 'ms/step %.3f'%d['ms_per_step'], 'sum %.3f'%sum(s.values()), ' '.join('%s=%.3f'%(k[:12],v) for k,v in s.items()))"; }
{ for lib in "" gpurun_variants/lib_async.so; do
    if [ -z "$lib" ]; then name=default; unset S360_LIB; else name=$lib; export S360_LIB=$PWD/$lib; fi
    timeout -s KILL 300 python bench.py --steps 60 --warmup 5 --no-e2e --no-cpu --no-cube6 --no-graph 2>/dev/null | line $name
  done
  S360_LIB=$PWD/gpurun_variants/lib_async.so timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -q 2>&1 | tail -2
  unset S360_LIB
  for lib in "" gpurun_variants/lib_mvbwd3.so; do
    if [ -z "$lib" ]; then unset S360_LIB; else export S360_LIB=$PWD/$lib; fi
    timeout -s KILL 300 python tools/views_bench.py 256 --no-separate 2>&1 | tail -1
  done
  unset S360_LIB
} > gpurun_out/r02f_ab.log 2>&1
python bench.py --config 4 --steps 20 --warmup 5 > gpurun_out/r02f_c4_n1.json 2> gpurun_out/r02f_c4_n1.err
python bench.py --config 5 --steps 10 --warmup 3 > gpurun_out/r02f_c5_n1.json 2> gpurun_out/r02f_c5_n1.err
cat gpurun_out/r02f_pytest.log gpurun_out/r02f_ab.log; tail -n 3 gpurun_out/r02f_c4_n1.err gpurun_out/r02f_c5_n1.err; head -c 700 gpurun_out/r02f_c4_n1.json; echo; head -c 1500 gpurun_out/r02f_c5_n1.json
