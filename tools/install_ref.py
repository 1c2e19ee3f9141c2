#!/usr/bin/env python
"""Install the UNMODIFIED reference into git-ignored ``baseline/_ref/`` so that it ships to the
GPU box with ``gpurun`` (``/root/reference`` does not exist there).

The reference is a flat directory of Python modules with no setup.py / pyproject, so
``pip install --target baseline/_ref /root/reference`` has nothing to build; the install is a
byte-for-byte copy of its ``*.py`` files (checked: a SHA-256 manifest is written next to them
and ``verify()`` re-checks it, so a test can prove the copy on the GPU box is the original).
Nothing under ``baseline/_ref`` is tracked by git (``.gitignore``) and nothing in the product
package reads it; it is used by ``tests/test_reference_step_gpu.py``, ``bench.py --mode step``
and ``tools/refstep.py``.

    python tools/install_ref.py            # copy (only in the build container)
    python tools/install_ref.py --verify   # re-hash the copy against its manifest
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")
MANIFEST = os.path.join(DST, "MANIFEST.sha256.json")


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def install(src: str = SRC, dst: str = DST) -> int:
    """Copy every .py of the reference (directory structure kept).  Returns the file count;
    0 when the reference tree is not present (GPU box: the shipped copy is used)."""
    if not os.path.isdir(src):
        return 0
    manifest = {}
    for d, _dirs, files in os.walk(src):
        for fn in files:
            if not fn.endswith(".py"):
                continue
            rel = os.path.relpath(os.path.join(d, fn), src)
            out = os.path.join(dst, rel)
            os.makedirs(os.path.dirname(out), exist_ok=True)
            shutil.copyfile(os.path.join(d, fn), out)
            manifest[rel] = _sha(out)
    with open(os.path.join(dst, "MANIFEST.sha256.json"), "w") as f:
        json.dump(manifest, f, indent=0, sort_keys=True)
    return len(manifest)


def installed(dst: str = DST) -> bool:
    return os.path.exists(os.path.join(dst, "train_step_final.py")) and os.path.exists(os.path.join(dst, "MANIFEST.sha256.json"))


def verify(dst: str = DST) -> bool:
    """True when every installed file still hashes to its manifest entry (i.e. is unmodified)."""
    with open(os.path.join(dst, "MANIFEST.sha256.json")) as f:
        manifest = json.load(f)
    return all(os.path.exists(os.path.join(dst, rel)) and _sha(os.path.join(dst, rel)) == h for rel, h in manifest.items())


if __name__ == "__main__":
    if "--verify" in sys.argv:
        ok = installed() and verify()
        print("baseline/_ref:", "unmodified" if ok else "MISSING or MODIFIED")
        sys.exit(0 if ok else 1)
    n = install()
    print(f"installed {n} reference files into {DST}" if n else f"{SRC} not present; nothing installed")
