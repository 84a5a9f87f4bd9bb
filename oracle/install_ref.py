"""Puts an UNMODIFIED copy of the reference tree under baseline/_ref/ (git-ignored, NOT gpurun-ignored) so that the
reference's own Python can be timed on the GPU box, where /root/reference does not exist.

    python -m oracle.install_ref

TEST / BENCH INFRASTRUCTURE ONLY (bench.py --impl reference and its cpu_baseline leg).  `pip install --target
baseline/_ref /root/reference` is not possible: the reference is a flat collection of scripts with neither setup.py
nor pyproject.toml ("Directory '/root/reference' is not installable"), so the tree is copied file by file instead.
Nothing under baseline/_ref is tracked by git and nothing in cdnet_b200/ imports it.
"""
import os
import shutil
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("CDNET_REF", "/root/reference")
DST = os.path.join(REPO, "baseline", "_ref")


def install(force=False):
    """returns the installed path, or None where the reference is not mounted (the GPU box uses what travelled)"""
    if not os.path.isfile(os.path.join(SRC, "postproc_other.py")):
        return DST if os.path.isfile(os.path.join(DST, "postproc_other.py")) else None
    if os.path.isdir(DST) and not force:
        same = all(os.path.exists(os.path.join(DST, f)) and
                   os.path.getmtime(os.path.join(DST, f)) >= os.path.getmtime(os.path.join(SRC, f))
                   for f in ("postproc_other.py", "test_dam.py", "my_transforms_direction.py"))
        if same:
            return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns(".git", "__pycache__", "*.pyc"))
    return DST


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
