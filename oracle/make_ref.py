"""Recipe for `baseline/_ref/`: the UNMODIFIED reference files of the hot path, for the CPU reference arm of bench.py.

The reference is a script repository without packaging metadata (no setup.py / pyproject.toml), so `pip install --target
baseline/_ref /root/reference` has nothing to install; this script does what that install would have done for the path: it COPIES,
byte for byte, the eight files `TrainModule.forward` executes and `models/arch/NBC2.py` from /root/reference into baseline/_ref/ — a directory
that is git-ignored (no reference source enters the history) but travels to the GPU box with the snapshot — and writes their SHA-256
next to them (MANIFEST.json).  `__graft_entry__.build()` runs it when /root/reference exists (in the build container); on the GPU box
the already-copied files are used.  TEST / BASELINE INFRASTRUCTURE: only bench.py's `--impl reference` arm and its `cpu_baseline` leg
import from baseline/_ref; the product package never does.

    models/arch/SpatialNet.py                       the network                       (§8 a4-a9)
    models/arch/base/{norm,non_linear,linear_group}.py   its norm / activation / LinearGroup modules
    models/io/{stft,norm}.py                        STFT / iSTFT, Norm                (§8 a1, a2, a10, a11)
    models/__init__.py, models/io/__init__.py       (empty package markers)

What cannot come along: SharedTrainer.py (needs pytorch_lightning / jsonargparse, absent from the image) and models/io/loss.py (needs
torchmetrics) — bench.py restates the 18 lines of TrainModule.forward (SharedTrainer.py:104-132, no arithmetic of its own) around the
copied modules and uses the oracle's torchmetrics restatement for the loss.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = [
    "models/__init__.py",
    "models/io/__init__.py",
    "models/arch/SpatialNet.py",
    "models/arch/base/norm.py",
    "models/arch/base/non_linear.py",
    "models/arch/base/linear_group.py",
    "models/io/stft.py",
    "models/io/norm.py",
    "models/arch/NBC2.py",  # BASELINE configs[3] (§8 a14): only tests/test_gpu_nbc2.py compares against it
]


def make_ref(src: str = "/root/reference", dst: str = os.path.join(ROOT, "baseline", "_ref")) -> bool:
    """Returns True when baseline/_ref holds the files (copied now, or already there and `src` is absent)."""
    if not os.path.isdir(os.path.join(src, "models")):
        return os.path.exists(os.path.join(dst, "MANIFEST.json"))
    manifest = {}
    for rel in FILES:
        s, d = os.path.join(src, rel), os.path.join(dst, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        manifest[rel] = hashlib.sha256(open(d, "rb").read()).hexdigest()
    with open(os.path.join(dst, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "what": "unmodified copies (sha256 of each file)", "files": manifest}, f, indent=1)
    return True


if __name__ == "__main__":
    ok = make_ref(*sys.argv[1:3])
    print("baseline/_ref:", "ready" if ok else "unavailable (no /root/reference and no earlier copy)")
