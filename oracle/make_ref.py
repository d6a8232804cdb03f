#!/usr/bin/env python
"""Install the UNMODIFIED reference into baseline/_ref (git-ignored, travels to the GPU box with the
snapshot): `pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy>`.

TEST / BENCH INFRASTRUCTURE.  The install is made from a scratch copy under /tmp because
/root/reference is read-only and setuptools writes egg-info/build directories into the source tree.
`--no-deps`: gymnasium / treevalue / jsonargparse / ... are not installable offline; the stand-ins
under oracle/refstubs replace them at run time (oracle/run_reference.py).  Recorded outcome:
DESIGN.md §6.  No reference source file is copied into the tracked tree.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_SRC = "/root/reference"
TARGET = os.path.join(ROOT, "baseline", "_ref")


def installed():
    return os.path.isfile(os.path.join(TARGET, "openrl", "__init__.py"))


def main(force=False):
    if installed() and not force:
        print("baseline/_ref already holds the reference install")
        return True
    if not os.path.isdir(REF_SRC):
        print("no /root/reference here (GPU box): using the prebuilt baseline/_ref" if installed() else "reference source absent")
        return installed()
    tmp = tempfile.mkdtemp(prefix="openrl_ref_")
    try:
        src = os.path.join(tmp, "reference")
        shutil.copytree(REF_SRC, src, ignore=shutil.ignore_patterns(".git"))
        if os.path.isdir(TARGET):
            shutil.rmtree(TARGET)
        os.makedirs(os.path.dirname(TARGET), exist_ok=True)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links", "/opt/wheelhouse",
               "--target", TARGET, src]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            print(r.stdout[-2000:], r.stderr[-2000:])
            return False
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    print("installed the reference into", TARGET)
    return installed()


if __name__ == "__main__":
    sys.exit(0 if main(force="--force" in sys.argv) else 1)
