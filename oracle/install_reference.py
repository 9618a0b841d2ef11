"""Place byte-identical copies of the reference's hot-path files under ``oracle/_ref/`` (TEST INFRASTRUCTURE).

The reference is a Python source tree without ``setup.py`` / ``pyproject.toml``, so ``pip install --target ...
/root/reference`` has nothing to install; this recipe is its equivalent for the model file of the hot path
(``src/models/clipcap.py``) and the executor that calls it (``src/trainers/clipcap_exector.py`` with its two base-class
modules), which ``tests/test_reference_executor.py`` drives, unmodified, against ``ClipCaptionPrefixB200`` on the GPU box.  ``oracle/_ref/`` is git-ignored (no reference source enters the history) but not
gpurun-ignored, so ``bench.py --impl reference`` and the in-line ``cpu_baseline`` can time the REAL, unmodified reference
module on the GPU box's host cores (``cpu_baseline.kind = "reference"``); without it they fall back to the pinned oracle
port (``kind = "port"``).  Runs only where ``/root/reference`` exists; ``__graft_entry__.build()`` calls it.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("EAVQA_REFERENCE", "/root/reference")
FILES = ["src/models/clipcap.py", "src/trainers/clipcap_exector.py", "src/trainers/base_executor.py",
         "src/trainers/metrics_processors.py"]


def install(verbose: bool = True) -> bool:
    if not os.path.isdir(REF):
        return False
    dst = os.path.join(ROOT, "oracle", "_ref")
    os.makedirs(dst, exist_ok=True)
    manifest = {}
    for rel in FILES:
        src = os.path.join(REF, rel)
        sub = os.path.join(dst, "trainers") if "/trainers/" in rel else dst
        os.makedirs(sub, exist_ok=True)
        out = os.path.join(sub, os.path.basename(rel))
        shutil.copyfile(src, out)
        with open(out, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(dst, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF, "sha256": manifest}, f, indent=1)
    if verbose:
        print("reference model files copied to", dst)
    return True


if __name__ == "__main__":
    install()
