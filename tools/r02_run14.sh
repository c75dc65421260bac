#!/bin/bash
# Programmatic dependent launch A/B: whole GPU suite with PDL on, bench configs 3/4/5 with S360_PDL=1 and =0.
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q 2>&1 | tail -n 12 > gpurun_out/r02m_pytest.log
for pdl in 1 0; do
  S360_PDL=$pdl timeout -s KILL 600 python bench.py --steps 200 --warmup 5 --no-cube6 > gpurun_out/r02m_c3_pdl$pdl.json 2> gpurun_out/r02m_c3_pdl$pdl.err
  S360_PDL=$pdl timeout -s KILL 600 python bench.py --config 4 --steps 50 --warmup 5 > gpurun_out/r02m_c4_pdl$pdl.json 2> gpurun_out/r02m_c4_pdl$pdl.err
  S360_PDL=$pdl timeout -s KILL 600 python bench.py --config 5 --steps 10 --warmup 3 --no-e2e > gpurun_out/r02m_c5_pdl$pdl.json 2> gpurun_out/r02m_c5_pdl$pdl.err
done
S360_PDL=1 timeout -s KILL 600 python bench.py --steps 200 --warmup 5 --no-cube6 --no-graph > gpurun_out/r02m_c3_eager_pdl1.json 2>/dev/null
S360_PDL=0 timeout -s KILL 600 python bench.py --steps 200 --warmup 5 --no-cube6 --no-graph > gpurun_out/r02m_c3_eager_pdl0.json 2>/dev/null
tail -n 6 gpurun_out/r02m_pytest.log
for f in gpurun_out/r02m_c*_pdl*.json; do echo $f; python -c "
import json
try:
  d=json.loads(open('$f').read().strip().splitlines()[-1])
  print({k:d.get(k) for k in ('value','ms_per_step')}, d.get('parity',{}).get('ok') if d.get('parity') else None)
except Exception as e: print('ERR', e)
"; done
tail -n 3 gpurun_out/r02m_c*_pdl1.err
