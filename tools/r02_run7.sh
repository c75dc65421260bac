#!/bin/bash
# Round-2 GPU session 7: GPU suite, bench, batched-path bench, config 4, ncu launch list + full capture of one step.
mkdir -p gpurun_out
timeout -s KILL 2400 python -m pytest tests -m gpu -q -rs 2>&1 > gpurun_out/r02g_pytest_full.log; tail -12 gpurun_out/r02g_pytest_full.log > gpurun_out/r02g_pytest.log
timeout -s KILL 900 python bench.py --steps 100 --warmup 10 > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err
timeout -s KILL 300 python bench.py --steps 100 --warmup 10 --no-graph --no-e2e --no-cpu --no-cube6 > gpurun_out/r02g_bench_nograph.json 2>> gpurun_out/r02g_bench.err
timeout -s KILL 300 python tools/views_bench.py 256 > gpurun_out/r02g_views.json 2>&1
python bench.py --config 4 --steps 30 --warmup 5 > gpurun_out/r02g_c4_n1.json 2> gpurun_out/r02g_c4_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'s360|render_|preprocess_|rs_|scan_|emit_|tile_|mb_|mse_' --launch-skip 60 --launch-count 90 --csv --log-file gpurun_out/r02g_launches.csv python bench.py --steps 6 --warmup 4 --no-e2e --no-cpu --no-cube6 --no-graph > gpurun_out/r02g_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'s360|render_|preprocess_|rs_|scan_|emit_|tile_|mb_' --launch-skip 60 --launch-count 15 -o gpurun_out/r02g_full -f python bench.py --steps 3 --warmup 4 --no-e2e --no-cpu --no-cube6 --no-graph > gpurun_out/r02g_ncu_full.log 2>&1
cat gpurun_out/r02g_pytest.log; tail -n 3 gpurun_out/r02g_bench.err gpurun_out/r02g_c4_n1.err; python - <<'PY'
import json
for f in ("r02g_bench", "r02g_bench_nograph", "r02g_c4_n1"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "ms/step", round(d["ms_per_step"], 4), "value", d["value"], "stages", {k: round(v, 4) for k, v in d["roofline"]["stages_ms"].items()})
        if d.get("parity"): print("  parity ok", d["parity"]["ok"], {k: d["parity"][k] for k in ("color", "d_means", "d_cov", "d_opac", "d_shs")})
        if d.get("e2e"): print("  e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"])
    except Exception as e:
        print(f, "ERR", e)
print(open("gpurun_out/r02g_views.json").read()[-900:])
PY
