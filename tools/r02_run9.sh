#!/bin/bash
# Round-2 final 1-GPU session: GPU suite, headline bench + reference arm, all BASELINE configs, batched path, configs 4/5,
# compute-sanitizer over the small cases.
mkdir -p gpurun_out
timeout -s KILL 2400 python -m pytest tests -m gpu -q -rs 2>&1 > gpurun_out/r02h_pytest_full.log; tail -12 gpurun_out/r02h_pytest_full.log > gpurun_out/r02h_pytest.log
timeout -s KILL 900 python bench.py --steps 200 --warmup 10 > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err
timeout -s KILL 900 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r02h_bench_reference.json 2>> gpurun_out/r02h_bench.err
timeout -s KILL 300 python tools/views_bench.py 256 > gpurun_out/r02h_views.json 2>&1
python bench.py --config 4 --steps 30 --warmup 5 > gpurun_out/r02h_c4_n1.json 2> gpurun_out/r02h_c4_n1.err
python bench.py --config 5 --steps 10 --warmup 3 > gpurun_out/r02h_c5_n1.json 2> gpurun_out/r02h_c5_n1.err
timeout -s KILL 900 python tests/tools/run_configs.py > gpurun_out/r02h_configs.log 2>&1; cp gpurun_out/configs.json gpurun_out/r02h_configs.json 2>/dev/null
timeout -s KILL 300 python tools/profile_config4.py > gpurun_out/r02h_profile_config4.log 2>&1
bash tools/gpu_sanitize.sh > gpurun_out/r02h_sanitize.log 2>&1
cat gpurun_out/r02h_pytest.log; tail -n 3 gpurun_out/r02h_bench.err gpurun_out/r02h_c4_n1.err gpurun_out/r02h_c5_n1.err; python - <<'PY'
import json
for f in ("r02h_bench", "r02h_bench_reference", "r02h_c4_n1", "r02h_c5_n1"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "ms/step", round(d["ms_per_step"], 4), "value", d["value"], "e2e", d.get("e2e") and (d["e2e"].get("ms_per_step"), d["e2e"]["value"]))
    except Exception as e:
        print(f, "ERR", e)
PY
head -2 gpurun_out/r02h_profile_config4.log; tail -5 gpurun_out/r02h_configs.log | cut -c1-400; cat gpurun_out/r02h_sanitize.log | tail -20
