#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 2400 python -m pytest tests -m gpu -q -rs 2>&1 | tail -8 > gpurun_out/r02i_pytest.log
python __graft_entry__.py --smoke > gpurun_out/r02i_smoke.log 2>&1
python bench.py --config 5 --steps 10 --warmup 3 > gpurun_out/r02i_c5_n1.json 2> gpurun_out/r02i_c5_n1.err
timeout -s KILL 900 python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu --no-cube6 > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err
cat gpurun_out/r02i_pytest.log; tail -n 4 gpurun_out/r02i_smoke.log; tail -n 3 gpurun_out/r02i_c5_n1.err gpurun_out/r02i_bench.err; python - <<'PY'
import json
for f in ("r02i_c5_n1", "r02i_bench"):
    d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, "ms/step", round(d["ms_per_step"], 4), {k: round(v, 4) for k, v in d["roofline"]["stages_ms"].items()})
PY
