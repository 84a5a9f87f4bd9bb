"""Builds the SIMT-emulated test double of libcdnet_b200.so with g++ (TEST INFRASTRUCTURE ONLY).

    python tests/simt/build.py [--force] [--asan]

--asan builds tests/simt/_build/asan/libcdnet_b200_simt_asan.so with -fsanitize=address and poisoned gaps
between the workspace slices; tests/simt/run_asan.sh runs the emulated kernel tests against it.

The kernel sources cdnet_b200/csrc/*.cu are compiled UNMODIFIED as C++ against tests/simt/include
(shims of cuda_runtime.h / cuda_fp16.h) and linked with tests/simt/simt_runtime.cpp into
tests/simt/_build/libcdnet_b200_simt.so, which exports the same C ABI (include/cdnet_b200.h) with
"device pointers" that are host pointers.  Only tests/ loads it.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(REPO, "cdnet_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libcdnet_b200_simt.so")
CXX = os.environ.get("CXX", "g++")
# -ffp-contract=off mirrors nvcc -fmad=false; -fno-strict-aliasing because the kernels type-pun through pointers
FLAGS = ["-std=c++17", "-O1", "-g", "-fPIC", "-pthread", "-ffp-contract=off", "-fno-strict-aliasing", "-w",
         "-I", os.path.join(HERE, "include")]


def _deps():
    return (glob.glob(os.path.join(CSRC, "*")) + glob.glob(os.path.join(HERE, "include", "*")) +
            [os.path.join(HERE, "simt_runtime.cpp"), os.path.join(REPO, "include", "cdnet_b200.h"), __file__])


def build(force=False, asan=False):
    OUT = os.path.join(HERE, "_build", "asan") if asan else os.path.join(HERE, "_build")
    LIB = os.path.join(OUT, "libcdnet_b200_simt_asan.so" if asan else "libcdnet_b200_simt.so")
    FLAGS = globals()["FLAGS"] + (["-fsanitize=address", "-fno-omit-frame-pointer", "-DCDNET_SIMT_ASAN"] if asan else [])
    CXX = globals()["CXX"]
    if asan and subprocess.run([CXX, "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip() == "libasan.so":
        CXX = "/usr/bin/g++"  # a compiler wrapper that cannot locate the sanitizer runtime: use the system one
    if (not force and os.path.exists(LIB)
            and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in _deps())):
        return LIB
    os.makedirs(OUT, exist_ok=True)
    jobs = []
    objs = []
    for src in sorted(glob.glob(os.path.join(CSRC, "*.cu"))) + [os.path.join(HERE, "simt_runtime.cpp")]:
        obj = os.path.join(OUT, os.path.basename(src).rsplit(".", 1)[0] + ".o")
        objs.append(obj)
        cmd = [CXX] + FLAGS + ["-x", "c++", "-c", src, "-o", obj]
        jobs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for cmd, pr in jobs:
        out = pr.communicate()[0].decode()
        if pr.returncode != 0:
            raise RuntimeError("g++ failed: %s\n%s" % (" ".join(cmd), out[-6000:]))
    tmp = LIB + ".tmp%d" % os.getpid()
    subprocess.check_call([CXX, "-shared", "-pthread", "-o", tmp] + objs + (["-fsanitize=address"] if asan else []))
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, asan="--asan" in sys.argv))
