"""bench.py contract, CPU side: the reference arm prints exactly ONE JSON line on stdout with the agreed keys (the GPU arm
shares the emit path and the key set is checked on its committed output in profiles/)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
        "data", "config", "e2e"}


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and KEYS <= set(d) and {"cores", "kind", "sample", "value"} <= set(d["cpu_baseline"])
    assert d["cpu_baseline"]["kind"] == "port" and d["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] == 0


def test_committed_gpu_bench_line_has_the_contract_keys():
    d = json.load(open(os.path.join(ROOT, "profiles", "r01_bench_final_b32.json")))
    assert KEYS | {"gpu_launches", "roofline", "cpu_baseline", "clocks"} <= set(d)
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"])
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"]) and d["e2e"]["h2d_bytes_per_step"] > 0
    assert d["config"]["workload"].startswith("SpatialNet-small 6ch F=129 T=250") and d["n_gpus"] == 1
