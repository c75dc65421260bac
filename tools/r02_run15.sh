#!/bin/bash
# 8-GPU session: per-rank PCIe ceiling (N=8 and N=1), config 3 (NUMA-bound pinned buffers) and config 4 (graphed decoder step).
N=${1:-8}
mkdir -p gpurun_out
tr() { timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) "$@"; }
tr tools/pcie_ceiling.py > gpurun_out/r02n_pcie_n$N.json 2> gpurun_out/r02n_pcie_n$N.err
timeout -s KILL 300 python tools/pcie_ceiling.py > gpurun_out/r02n_pcie_n1.json 2> gpurun_out/r02n_pcie_n1.err
tr bench.py --gpus $N --steps 20 --warmup 5 --no-cube6 > gpurun_out/r02n_c3_n$N.json 2> gpurun_out/r02n_c3_n$N.err
tr bench.py --gpus $N --config 4 --steps 20 --warmup 5 > gpurun_out/r02n_c4_n$N.json 2> gpurun_out/r02n_c4_n$N.err
cat gpurun_out/r02n_pcie_n$N.json gpurun_out/r02n_pcie_n1.json
for f in gpurun_out/r02n_c3_n$N.json gpurun_out/r02n_c4_n$N.json; do echo $f; python -c "
import json
try:
  d=json.loads(open('$f').read().strip().splitlines()[-1])
  print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e'] and {k:d['e2e'].get(k) for k in ('value','ms_per_step','numa_bind')})
except Exception as e: print('ERR', e)
"; done
tail -n 4 gpurun_out/r02n_pcie_n$N.err gpurun_out/r02n_c3_n$N.err gpurun_out/r02n_c4_n$N.err | grep -v OMP_NUM | grep -v "^\*\*\*"
