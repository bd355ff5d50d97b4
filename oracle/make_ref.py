#!/usr/bin/env python
"""Stage the UNMODIFIED reference model code into ``oracle/_ref/`` so that it travels to the GPU box.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (like everything under ``oracle/``).

``/root/reference`` exists only in the build container; the GPU box receives the working tree (minus ``.git``).
``oracle/_ref/`` is git-ignored -- reference sources never enter this repository's history -- but it is NOT
gpurun-ignored, so a staged copy rides along with the snapshot exactly like the built ``.so`` files do.  This
recipe is what ``__graft_entry__.build()`` runs when ``/root/reference`` is present.

What is staged: the importable model package and its one top-level dependency, byte for byte
(``models/*.py`` of the live model: ``cas_mvsnet.py``, ``module.py``, ``homography.py``, ``geometry.py``, ``FMT.py``,
``position_encoding.py``, ``__init__.py``; ``utils.py``).  Nothing is edited: the one load-time patch the reference
needs below 576x1019 (debug prints at models/cas_mvsnet.py:275-286) is applied in memory by ``oracle/ref_loader.py``.
A ``MANIFEST.json`` records the sha256 of every file so the loader can verify what it runs.

Consumers: ``bench.py --impl reference`` (the reference's own DepthNet on the host cores), ``bench.py``'s
``incumbent_gpu`` / ``full_forward`` legs (the reference's PyTorch path on the B200), ``tests/`` (full-size parity
cross-checks).  Never the product path.

Usage:  python oracle/make_ref.py [--src /root/reference]
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = (
    "utils.py",
    "models/__init__.py",
    "models/cas_mvsnet.py",
    "models/module.py",
    "models/homography.py",
    "models/geometry.py",
    "models/FMT.py",
    "models/position_encoding.py",
)


def _sha(path: str) -> str:
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def stage(src: str = "/root/reference", dest: str = DEST) -> str | None:
    """Copy FILES from `src` to `dest`; returns dest, or None when `src` is absent (GPU box: use what travelled)."""
    if not os.path.isdir(src):
        return dest if os.path.exists(os.path.join(dest, "MANIFEST.json")) else None
    manifest = {}
    for rel in FILES:
        s, d = os.path.join(src, rel), os.path.join(dest, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if not (os.path.exists(d) and _sha(d) == _sha(s)):
            shutil.copyfile(s, d)
        manifest[rel] = _sha(d)
    with open(os.path.join(dest, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "files": manifest}, f, indent=1, sort_keys=True)
    return dest


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    out = stage(ap.parse_args().src)
    print(out if out else "reference checkout not found and nothing staged", file=sys.stderr if out is None else sys.stdout)
    sys.exit(0 if out else 1)
