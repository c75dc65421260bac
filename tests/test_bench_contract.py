"""bench.py's output contract, checked without a GPU: the reference arm really runs here (the CPU oracle port on the full
config-3 workload, one step), and the committed GPU line of the closing session (profiles/r02_bench.json) carries every key
the driver reads and is consistent with itself."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "cpu_baseline", "e2e")


def test_reference_arm_prints_one_contract_line_on_the_cpu():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3"],
                       cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in BASE + ("impl",):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "gaussians_per_s_fwd_bwd" and d["unit"] == "Gaussians/s"
    assert d["steps"] == 1 and d["warmup"] >= 3 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["config"]["P"] == 1048576 and d["config"]["image"] == [512, 1024] and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "1048576" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert abs(d["value"] - 1048576 / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]


def test_committed_gpu_line_has_every_contract_key_and_is_self_consistent():
    d = json.loads(open(os.path.join(ROOT, "profiles", "r02_bench.json")).read().strip().splitlines()[-1])
    for k in BASE + ("roofline", "gpu_launches", "clocks", "parity"):
        assert k in d, k
    assert d["metric"] == "gaussians_per_s_fwd_bwd" and d["dtype"] == "f32" and d["data"] == "synthetic" and d["n_gpus"] == 1
    assert d["warmup"] >= 3 and d["scaling"] == "weak" and d["vs_baseline"] is None and "workload" in d["config"]
    assert "model" not in d["config"]
    assert abs(d["value"] - d["config"]["P"] / (d["ms_per_step"] * 1e-3)) <= 1e-3 * d["value"]
    rf = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in rf, k
    assert rf["bound"] in ("hbm", "tensor") and rf["unit"] in ("GB/s", "TFLOP/s")
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-6 and 0 < rf["frac"] < 1
    cb = d["cpu_baseline"]
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in cb, k
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 300e6 and e["d2h_bytes_per_step"] >= 4 and 0 < e["value"] < d["value"]
    assert d["gpu_launches"] > 0 and d["clocks"]["sm_mhz"] > 0 and isinstance(d["clocks"]["reasons"], list)
    assert d["parity"]["ok"] is True and d["parity"]["tolerance"] == 1e-4
