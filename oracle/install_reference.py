"""Recipe that installs the UNMODIFIED reference (`/root/reference`, pure Python) under `oracle/_ref/` so that the CPU arm
of `bench.py` can run the reference's own `PLS.calculate_particle_update` on the GPU box, where `/root/reference` does
not exist.  TEST / MEASUREMENT INFRASTRUCTURE: nothing in the product package imports it.

    python oracle/install_reference.py        (also run by __graft_entry__.build() when /root/reference is present)

`pip install --no-index --no-build-isolation --no-deps --target oracle/_ref/src <copy of /root/reference>`: the reference's
pyproject has no package layout of its own (`[tool.uv] package = false`); setuptools' src-layout discovery installs the
modules found under its `src/` directory flat, and the reference's modules import one another as `src.<module>`
(src/projected_langevin_sampling/projected_langevin_sampling.py:3), so the target directory is itself named `src` and
`oracle/_ref` goes on sys.path (an implicit namespace package).  `oracle/_ref/` is git-ignored (no reference source ever
enters the history) but not gpurun-ignored, so the installed copy travels to the GPU box like the built `.so`.  The
reference needs gpytorch, which is absent from this image: it is run behind `oracle/gpytorch_stub`.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference"
REF_DST = os.path.join(HERE, "_ref")
MARKER = os.path.join(REF_DST, "src", "projected_langevin_sampling", "projected_langevin_sampling.py")


def installed() -> bool:
    return os.path.exists(MARKER)


def install(force: bool = False) -> str:
    """Returns 'installed', 'present' or 'unavailable: <why>'."""
    if installed() and not force:
        return "present"
    if not os.path.isdir(REF_SRC):
        return "unavailable: /root/reference does not exist on this machine"
    tmp = tempfile.mkdtemp(prefix="pls_ref_")
    try:
        work = os.path.join(tmp, "reference")
        shutil.copytree(REF_SRC, work)  # /root/reference is read-only and the build writes an egg-info next to the sources
        if os.path.isdir(REF_DST):
            shutil.rmtree(REF_DST)
        os.makedirs(REF_DST)
        cmd = [sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--target", os.path.join(REF_DST, "src"), work]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0 or not installed():
            return "unavailable: pip install failed: " + (r.stderr.strip().splitlines() or ["?"])[-1]
        return "installed"
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
