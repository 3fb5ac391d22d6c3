#!/usr/bin/env python
"""Populate the git-ignored baseline/_ref/ with the UNMODIFIED reference package so that `bench.py --impl reference` can time
the reference's own Python (`reference_python` leg) on the GPU box, where /root/reference does not exist (baseline/_ref is
git-ignored but travels with gpurun).  Run where /root/reference exists (the build container; `__graft_entry__.build()` calls it).

First choice is the install the build contract names (pip, offline, from a /tmp copy because the build writes an egg-info into
the source tree, --no-deps because `simulator` / `lightning` are not in the wheelhouse); if that fails the package directory
`src/alphazero_implementation` is copied as it is.  Either way the two missing third-party packages come from oracle/shims at
run time (a CPU restatement of `simulator.game.connect`, an nn.Module-based `lightning`): nothing here is product source.
"""
from __future__ import annotations

import json
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference"
DEST = os.path.join(ROOT, "baseline", "_ref")


def install() -> dict:
    if not os.path.isdir(os.path.join(REFERENCE, "src", "alphazero_implementation")):
        return {"installed": False, "why": f"{REFERENCE} not present"}
    shutil.rmtree(DEST, ignore_errors=True)
    os.makedirs(DEST, exist_ok=True)
    how = None
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "reference")
        shutil.copytree(REFERENCE, src, ignore=shutil.ignore_patterns(".git", "*.ipynb", "lightning_logs"))
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links", "/opt/wheelhouse",
               "--target", DEST, src]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode == 0 and os.path.isdir(os.path.join(DEST, "alphazero_implementation")):
            how = "pip install --no-index --no-build-isolation --no-deps --target baseline/_ref (from a /tmp copy)"
        else:
            shutil.rmtree(DEST, ignore_errors=True)
            os.makedirs(DEST, exist_ok=True)
            shutil.copytree(os.path.join(REFERENCE, "src", "alphazero_implementation"), os.path.join(DEST, "alphazero_implementation"))
            how = "copied src/alphazero_implementation (pip failed: " + (proc.stderr.strip().splitlines() or ["?"])[-1][:200] + ")"
    info = {"installed": True, "how": how}
    json.dump(info, open(os.path.join(DEST, "INSTALL.json"), "w"))
    return info


if __name__ == "__main__":
    print(json.dumps(install()))
