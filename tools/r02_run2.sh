#!/bin/bash
# Round-2 GPU session 2: full GPU suite (complete log), parity diagnosis, binning A/B, ncu of the new kernels, bench.
mkdir -p gpurun_out
timeout -s KILL 2400 python -m pytest tests -m gpu -q -rs 2>&1 > gpurun_out/r02b_pytest_full.log; tail -30 gpurun_out/r02b_pytest_full.log > gpurun_out/r02b_pytest.log
timeout -s KILL 600 python tests/tools/diag_parity.py > gpurun_out/r02b_diag_parity.json 2> gpurun_out/r02b_diag_parity.err
line() { python -c "
import sys,json
d=json.loads(sys.stdin.read()); s=d['roofline']['stages_ms']
print('$1', 'ms/step %.3f'%d['ms_per_step'], 'sum %.3f'%sum(s.values()), ' '.join('%s=%.3f'%(k[:12],v) for k,v in s.items()))"; }
{ S360_FORCE_RADIX_BINNING=1 timeout -s KILL 300 python bench.py --steps 60 --warmup 5 --no-e2e --no-cpu --no-cube6 2>/dev/null | line radix_binning
  timeout -s KILL 300 python bench.py --steps 60 --warmup 5 --no-e2e --no-cpu --no-cube6 2>/dev/null | line matrix_binning
  timeout -s KILL 300 python bench.py --steps 60 --warmup 5 --no-e2e --no-cpu --no-cube6 --exact-counts 2>/dev/null | line matrix_exact_counts
} > gpurun_out/r02b_ab.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'mb_|tile_scan|rs_onesweep|rs_global|preprocess_kernel' --launch-skip 40 --launch-count 12 -o gpurun_out/r02b_binning -f python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-cube6 > gpurun_out/r02b_ncu_binning.log 2>&1
timeout -s KILL 900 python bench.py --steps 100 --warmup 10 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
cat gpurun_out/r02b_pytest.log; cat gpurun_out/r02b_ab.log; tail -3 gpurun_out/r02b_bench.err; head -c 1500 gpurun_out/r02b_diag_parity.json
