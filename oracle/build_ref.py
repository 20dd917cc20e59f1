"""Install the UNMODIFIED reference package into ``oracle/_ref/`` so that it travels to the GPU box.

Test / benchmark infrastructure only (see oracle/__init__.py): ``bench.py --impl reference`` times the reference's own
``Xtractor`` / ``PLDA_scoring`` from here (``cpu_baseline.kind = "reference"``) and the ``reference``-marked tests run
next to the CUDA path.  ``oracle/_ref/`` is git-ignored (no reference source ever enters the history) but not
gpurun-ignored.  Run in the build container, where ``/root/reference`` exists:

    python oracle/build_ref.py

Recipe: ``pip install --no-index --no-deps --no-build-isolation --target oracle/_ref <copy of /root/reference>`` (from a
copy under /tmp: the reference tree is read-only).  The reference's ``setup.py`` lists no ``packages=``, so pip installs
only its scripts; the ``sidekit`` package directory is then installed by a plain tree copy, which is all a pure-Python
"install" is.  Nothing is patched: the stubs (h5py / matplotlib / soundfile) and the two patches P1 / P2 that its
``forward`` needs are applied at import time by ``oracle/ref_import.py``, exactly as in the build container.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = os.environ.get("SIDEKIT_REFERENCE_SRC", "/root/reference")


def build(verbose=True):
    if not os.path.isdir(os.path.join(SRC, "sidekit")):
        if verbose:
            print("oracle/build_ref: %s not present; keeping %s as it is" % (SRC, DEST))
        return os.path.isdir(os.path.join(DEST, "sidekit"))
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    os.makedirs(DEST)
    how = []
    with tempfile.TemporaryDirectory() as tmp:
        work = os.path.join(tmp, "reference")
        shutil.copytree(SRC, work, ignore=shutil.ignore_patterns(".git", "egs"))
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-deps", "--no-build-isolation", "--quiet",
               "--target", DEST, work]
        rc = subprocess.run(cmd, capture_output=True, text=True)
        how.append("pip rc=%d" % rc.returncode)
        if not os.path.isfile(os.path.join(DEST, "sidekit", "__init__.py")):
            shutil.copytree(os.path.join(work, "sidekit"), os.path.join(DEST, "sidekit"),
                            ignore=shutil.ignore_patterns("__pycache__"))
            how.append("package tree copied (setup.py declares no packages)")
    with open(os.path.join(DEST, "INSTALLED_FROM.txt"), "w") as f:
        f.write("source: %s\nhow: %s\n" % (SRC, "; ".join(how)))
    if verbose:
        print("oracle/build_ref: installed the reference into %s (%s)" % (DEST, "; ".join(how)))
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
