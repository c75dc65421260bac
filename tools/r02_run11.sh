#!/bin/bash
# Config-5 end-to-end session (N GPUs): multi-rank parity (incl. the pipelined upload+broadcast) and config 5 under torchrun,
# plus the same at N=1 for the efficiency denominator.
N=${1:-2}
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q -rs 2>&1 | tail -15 > gpurun_out/r02j_multirank_n$N.log
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) \
  bench.py --gpus $N --config 5 --steps 10 --warmup 3 > gpurun_out/r02j_c5_n$N.json 2> gpurun_out/r02j_c5_n$N.err
timeout -s KILL 900 python bench.py --config 5 --steps 10 --warmup 3 > gpurun_out/r02j_c5_n1.json 2> gpurun_out/r02j_c5_n1.err
cat gpurun_out/r02j_multirank_n$N.log
for f in gpurun_out/r02j_c5_n$N.json gpurun_out/r02j_c5_n1.json; do echo $f; python -c "
import json
try:
  d=json.loads(open('$f').read().strip().splitlines()[-1])
  print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e'] and {k:d['e2e'][k] for k in ('value','ms_per_step')})
except Exception as e: print('ERR', e)
"; done
tail -n 5 gpurun_out/r02j_c5_n$N.err
tail -n 5 gpurun_out/r02j_c5_n1.err
