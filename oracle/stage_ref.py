#!/usr/bin/env python3
"""Stage the UNMODIFIED reference package into ``oracle/_ref/`` so that it can be driven on the GPU box.

TEST / BASELINE INFRASTRUCTURE ONLY: nothing under ``toycrystals_b200/`` imports it.  ``bench.py --impl reference`` and
``bench.py``'s ``cpu_baseline`` leg import ``toycrystals`` from ``oracle/_ref`` (with ``oracle/_stubs`` in front for the
``matplotlib`` import at the reference's module scope) and time the reference's OWN
``sample_reverse_sde_euler_maruyama`` (sde_score_model.py:507-569) — on the host cores, and as eager PyTorch on the B200.

``/root/reference`` exists only in the build container; ``oracle/_ref/`` is git-ignored (the reference's sources never
enter this repository's history) but not gpurun-ignored, so the staged copy travels to the GPU box with the snapshot, as
``baseline/_ref`` would for an installable reference.

What it does: the offline equivalent of ``pip install --no-deps --target oracle/_ref /root/reference``.  The reference's
build backend (hatchling) is not in this image, so pip cannot build the wheel; for a pure-Python package with
``packages = ["src/toycrystals"]`` (pyproject.toml:27-28) the installed tree is exactly a copy of that directory, which is
what this script produces, with a MANIFEST of SHA-256 sums so the GPU-side run can state what it executed.

    python oracle/stage_ref.py            # (re)creates oracle/_ref/toycrystals + oracle/_ref/MANIFEST.json
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")


def stage(verbose: bool = True) -> bool:
    pkg = os.path.join(SRC, "src", "toycrystals")
    if not os.path.isdir(pkg):
        if verbose:
            print(f"stage_ref: {SRC} not present (GPU box?) - keeping whatever is in {DST}")
        return os.path.isdir(os.path.join(DST, "toycrystals"))
    how = "copy of src/toycrystals (hatchling missing: pip cannot build the wheel offline)"
    tmp = DST + ".tmp"
    shutil.rmtree(tmp, ignore_errors=True)
    r = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
                        "--find-links", "/opt/wheelhouse", "--target", tmp, SRC], capture_output=True, text=True)
    if r.returncode == 0 and os.path.isdir(os.path.join(tmp, "toycrystals")):
        how = "pip install --no-deps --target"
    else:
        shutil.rmtree(tmp, ignore_errors=True)
        os.makedirs(tmp)
        shutil.copytree(pkg, os.path.join(tmp, "toycrystals"), ignore=shutil.ignore_patterns("__pycache__"))
    man = {"how": how, "files": {}}
    for root, _, files in os.walk(os.path.join(tmp, "toycrystals")):
        for f in sorted(files):
            if f.endswith(".py"):
                p = os.path.join(root, f)
                man["files"][os.path.relpath(p, tmp)] = hashlib.sha256(open(p, "rb").read()).hexdigest()
    json.dump(man, open(os.path.join(tmp, "MANIFEST.json"), "w"), indent=1)
    shutil.rmtree(DST, ignore_errors=True)
    os.rename(tmp, DST)
    if verbose:
        print(f"stage_ref: {len(man['files'])} files -> {DST} ({how})")
    return True


def import_reference():
    """Import the staged reference's sampling module; returns (module, manifest) or raises ImportError."""
    if not os.path.isdir(os.path.join(DST, "toycrystals")):
        raise ImportError(f"{DST}/toycrystals missing: run `python oracle/stage_ref.py` in the build container")
    for p in (DST, os.path.join(HERE, "_stubs")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import toycrystals.models.sde_score_model as ref
    assert os.path.realpath(ref.__file__).startswith(os.path.realpath(DST)), ref.__file__
    man = json.load(open(os.path.join(DST, "MANIFEST.json")))
    return ref, man


if __name__ == "__main__":
    raise SystemExit(0 if stage() else 1)
