"""Install the UNMODIFIED reference package into baseline/_ref (git-ignored, but it travels to the GPU box).

    python scripts/install_reference.py [--force]

The recipe is the one the task contract prescribes: pip from a copy of /root/reference under /tmp (the source tree is
read-only and setuptools writes egg-info next to it), no index, no build isolation, no dependency resolution
(numpy / numba / pandas are already in the image).  `setuptools_scm` is absent, so the package reports version 0.0.0
(its own fallback, numpy_groupies/__init__.py:46-56).  Outcome recorded in DESIGN.md section 7.
Consumers: bench.py (--impl reference and cpu_baseline time aggregate_numpy / aggregate_numba from here) and
tests/test_reference_suite_*.py (the reference's own pytest suite with aggregate_cuda registered).
"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
DEST = os.path.join(ROOT, "baseline", "_ref")


def installed():
    return os.path.exists(os.path.join(DEST, "numpy_groupies", "aggregate_numpy.py"))


def install(force=False):
    if installed() and not force:
        return DEST
    if not os.path.isdir(REF):
        raise RuntimeError(f"{REF} is not present: baseline/_ref can only be (re)built in the build container")
    tmp = tempfile.mkdtemp(prefix="npg_ref_")
    try:
        src = os.path.join(tmp, "reference")
        shutil.copytree(REF, src)
        if os.path.isdir(DEST):
            shutil.rmtree(DEST)
        os.makedirs(DEST, exist_ok=True)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--target", DEST, src]
        res = subprocess.run(cmd, capture_output=True, text=True, cwd=tmp)
        if res.returncode != 0 or not installed():
            raise RuntimeError("pip install of the reference failed:\n" + res.stdout[-2000:] + res.stderr[-2000:])
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return DEST


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
