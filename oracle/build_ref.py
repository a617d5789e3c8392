"""Recipe for ``oracle/_ref``: the UNMODIFIED reference package, installed from where it lies.

    python oracle/build_ref.py            # run in the build container (needs /root/reference)

TEST INFRASTRUCTURE (see oracle/__init__.py).  ``oracle/_ref`` is git-ignored (outputs only, never part
of the history) but travels to the GPU box with the snapshot, where ``bench.py --impl reference`` and
the ``cpu_baseline`` leg time it as the CPU arm (``kind: "reference"``).

Steps:
  1. ``pip install --no-index --no-build-isolation --no-deps --target oracle/_ref`` of a scratch copy of
     ``/root/reference`` (setup.py + memento/; the tree itself is read-only, setuptools wants to write
     ``build/`` next to setup.py).  The installed ``memento/*.py`` are byte-identical to the reference
     (checked below).
  2. The reference imports four third-party modules it never uses on this path (``patsy``,
     ``statsmodels``, ``scanpy``, ``matplotlib``) and this image does not have them:
     the empty stub packages of ``tests/golden/ref_shim`` (ours) are copied next to it, on disk, so that
     joblib/loky worker processes can import them too.
"""
import filecmp
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REFERENCE = "/root/reference"
TARGET = os.path.join(HERE, "_ref")
SHIM = os.path.join(ROOT, "tests", "golden", "ref_shim")
FILES = ("__init__.py", "bootstrap.py", "estimator.py", "hypothesis_test.py", "main.py", "simulate.py", "util.py")


def up_to_date():
    return all(os.path.exists(os.path.join(TARGET, "memento", f)) and
               (not os.path.isdir(REFERENCE) or
                filecmp.cmp(os.path.join(TARGET, "memento", f), os.path.join(REFERENCE, "memento", f), shallow=False))
               for f in FILES) and os.path.exists(os.path.join(TARGET, "patsy", "__init__.py"))


def build(force=False):
    """Returns True when oracle/_ref is in place (built now or earlier), False when it cannot be built here."""
    if not force and up_to_date():
        return True
    if not os.path.isdir(os.path.join(REFERENCE, "memento")):
        return os.path.exists(os.path.join(TARGET, "memento", "main.py"))
    shutil.rmtree(TARGET, ignore_errors=True)
    with tempfile.TemporaryDirectory() as tmp:
        shutil.copy(os.path.join(REFERENCE, "setup.py"), tmp)
        shutil.copytree(os.path.join(REFERENCE, "memento"), os.path.join(tmp, "memento"))
        subprocess.run([sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation",
                        "--no-deps", "--find-links", "/opt/wheelhouse", "--target", TARGET, tmp], check=True)
    for name in os.listdir(SHIM):
        src = os.path.join(SHIM, name)
        if os.path.isdir(src):
            shutil.copytree(src, os.path.join(TARGET, name), dirs_exist_ok=True)
    for f in FILES:
        assert filecmp.cmp(os.path.join(TARGET, "memento", f), os.path.join(REFERENCE, "memento", f), shallow=False), f
    return True


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref:", "ready" if ok else "unavailable (no /root/reference here and no earlier build)")
