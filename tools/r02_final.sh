#!/bin/bash
# Round-2 closing 1-GPU session on the final code: GPU suite, smoke, headline bench + reference arm, configs 4 / 5, six-face
# side bench, ncu launch list + full capture of one eager step, compute-sanitizer over the small cases.
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q -rs 2>&1 > gpurun_out/r02z_pytest_full.log; tail -12 gpurun_out/r02z_pytest_full.log > gpurun_out/r02z_pytest.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02z_smoke.log 2>&1
timeout -s KILL 900 python bench.py --steps 200 --warmup 10 > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err
timeout -s KILL 900 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r02z_bench_reference.json 2>> gpurun_out/r02z_bench.err
timeout -s KILL 300 python bench.py --steps 200 --warmup 10 --no-graph --no-e2e --no-cpu --no-cube6 > gpurun_out/r02z_bench_nograph.json 2>> gpurun_out/r02z_bench.err
timeout -s KILL 300 python tools/views_bench.py 256 > gpurun_out/r02z_views.json 2>&1
timeout -s KILL 300 python bench.py --config 4 --steps 50 --warmup 5 > gpurun_out/r02z_c4_n1.json 2> gpurun_out/r02z_c4_n1.err
timeout -s KILL 300 python bench.py --config 5 --steps 10 --warmup 3 > gpurun_out/r02z_c5_n1.json 2> gpurun_out/r02z_c5_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'s360|render_|preprocess_|rs_|scan_|emit_|tile_|mb_|mse_' --launch-skip 60 --launch-count 90 --csv --log-file gpurun_out/r02z_launches.csv python bench.py --steps 6 --warmup 4 --no-e2e --no-cpu --no-cube6 --no-graph > gpurun_out/r02z_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'s360|render_|preprocess_|rs_|scan_|emit_|tile_|mb_' --launch-skip 60 --launch-count 15 -o gpurun_out/r02z_full -f python bench.py --steps 3 --warmup 4 --no-e2e --no-cpu --no-cube6 --no-graph > gpurun_out/r02z_ncu_full.log 2>&1
bash tools/gpu_sanitize.sh > gpurun_out/r02z_sanitize.log 2>&1
cat gpurun_out/r02z_pytest.log; cat gpurun_out/r02z_smoke.log | tail -n 4; tail -n 3 gpurun_out/r02z_bench.err gpurun_out/r02z_c4_n1.err gpurun_out/r02z_c5_n1.err; python - <<'PY'
import json
for f in ("r02z_bench", "r02z_bench_reference", "r02z_bench_nograph", "r02z_c4_n1", "r02z_c5_n1"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "ms/step", round(d["ms_per_step"], 4), "value", d["value"], "e2e", d.get("e2e") and (d["e2e"].get("ms_per_step"), d["e2e"]["value"]))
        if d.get("parity"): print("  parity ok", d["parity"]["ok"], {k: d["parity"][k] for k in ("color", "d_means", "d_cov", "d_opac", "d_shs")})
        if d.get("roofline") and d["roofline"].get("stages_ms"): print("  stages", {k: round(v, 4) for k, v in d["roofline"]["stages_ms"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
tail -c 600 gpurun_out/r02z_views.json; tail -n 12 gpurun_out/r02z_sanitize.log; ls -la gpurun_out/r02z_full.ncu-rep
