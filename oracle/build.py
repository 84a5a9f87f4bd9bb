"""Compile the oracle's C restatement (oracle/ws_flood.c) into oracle/_build/liboracle.so.

TEST INFRASTRUCTURE ONLY.  Called by `__graft_entry__.build()` and lazily by `oracle.clib`.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "liboracle.so")
SRC = [os.path.join(HERE, "ws_flood.c")]


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if (not force and os.path.exists(LIB)
            and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in SRC)):
        return LIB
    tmp = LIB + ".tmp%d" % os.getpid()
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", tmp] + SRC + ["-lm"])
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
