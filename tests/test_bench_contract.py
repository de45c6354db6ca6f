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
    # "reference" = the unmodified reference modules copied into baseline/_ref by oracle/make_ref.py (present wherever build() ran with
    # /root/reference in reach); "port" = the op-set restatement, the fallback
    have_ref = os.path.exists(os.path.join(ROOT, "baseline", "_ref", "MANIFEST.json"))
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port"), d["cpu_baseline"]
    assert d["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] == 0


def test_reference_arm_modules_are_unmodified_and_agree_with_the_port():
    """baseline/_ref (when present): every file still has the SHA-256 recorded when it was copied and — in the build container — is
    byte-identical to /root/reference; the copied modules + the restated TrainModule.forward compute what the op-set port computes."""
    import hashlib

    import pytest
    import torch

    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.exists(os.path.join(ref_dir, "MANIFEST.json")):
        pytest.skip("baseline/_ref not built (no /root/reference here)")
    man = json.load(open(os.path.join(ref_dir, "MANIFEST.json")))["files"]
    assert "models/arch/SpatialNet.py" in man and "models/io/stft.py" in man
    for rel, sha in man.items():
        data = open(os.path.join(ref_dir, rel), "rb").read()
        assert hashlib.sha256(data).hexdigest() == sha, rel
        if os.path.isdir("/root/reference/models"):
            assert data == open(os.path.join("/root/reference", rel), "rb").read(), rel
    sys.path.insert(0, ROOT)
    import bench
    from oracle import eager_gpu as E
    from oracle import spatialnet_oracle as O

    saved = {k: v for k, v in sys.modules.items() if k == "models" or k.startswith("models.")}
    try:
        for k in saved:
            del sys.modules[k]
        arch, fwd = bench.reference_modules(num_layers=2)
        x, _ = bench.synth_batch(1, 5)
        x = x[..., :128 * 20].contiguous()
        cfg = dict(O.SMALL_CFG, num_layers=2)
        with torch.no_grad():
            est_ref = fwd(x.clone())
            est_port = E.io_forward(O.synth_params(cfg, 2), x, cfg, 256, 128, 0)
        assert O.rel_l2(est_ref, est_port) < 1e-5
    finally:  # other tests import the full `models` package of the live reference: leave no partial copy behind
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
            del sys.modules[k]
        sys.modules.update(saved)
        if bench.REF_DIR in sys.path:
            sys.path.remove(bench.REF_DIR)


def test_committed_gpu_bench_line_has_the_contract_keys():
    d = json.load(open(os.path.join(ROOT, "profiles", "r01_bench_final_b32.json")))
    assert KEYS | {"gpu_launches", "roofline", "cpu_baseline", "clocks"} <= set(d)
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"])
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"]) and d["e2e"]["h2d_bytes_per_step"] > 0
    assert d["config"]["workload"].startswith("SpatialNet-small 6ch F=129 T=250") and d["n_gpus"] == 1
