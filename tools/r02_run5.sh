#!/bin/bash
# Round-2 multi-GPU session (N = number of visible GPUs): multi-rank parity test, configs 3/4/5 under torchrun.
N=${1:-2}
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q -rs 2>&1 | tail -15 > gpurun_out/r02e_multirank_n$N.log
run() { timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) bench.py --gpus $N "$@"; }
run --steps 20 --warmup 5 --no-cube6 > gpurun_out/r02e_c3_n$N.json 2> gpurun_out/r02e_c3_n$N.err
run --config 4 --steps 20 --warmup 5 > gpurun_out/r02e_c4_n$N.json 2> gpurun_out/r02e_c4_n$N.err
run --config 5 --steps 10 --warmup 3 > gpurun_out/r02e_c5_n$N.json 2> gpurun_out/r02e_c5_n$N.err
if [ "$2" == "single" ]; then
  python bench.py --config 4 --steps 20 --warmup 5 > gpurun_out/r02e_c4_n1.json 2> gpurun_out/r02e_c4_n1.err
  python bench.py --config 5 --steps 10 --warmup 3 > gpurun_out/r02e_c5_n1.json 2> gpurun_out/r02e_c5_n1.err
fi
cat gpurun_out/r02e_multirank_n$N.log
for f in gpurun_out/r02e_c*_n*.json; do echo $f; python -c "
import json,sys
try:
  d=json.loads(open('$f').read().strip().splitlines()[-1])
  print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e'] and {k:d['e2e'][k] for k in ('value','ms_per_step')})
except Exception as e: print('ERR', e)
"; done
tail -n 3 gpurun_out/r02e_c5_n$N.err gpurun_out/r02e_c4_n$N.err gpurun_out/r02e_c3_n$N.err
