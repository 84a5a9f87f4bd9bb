"""Builds cdnet_b200/libcdnet_b200_<name>.so with extra -D defines, for A/B runs of compile-time tunings on one GPU box:
    python tools/build_variant.py lab8 -DCDNET_LAB_ROWS=8
    CDNET_B200_LIB=cdnet_b200/libcdnet_b200_lab8.so python bench.py ..."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cdnet_b200 import build as B  # noqa: E402


def main():
    name, defs = sys.argv[1], sys.argv[2:]
    objdir = os.path.join(B.HERE, "build_" + name)
    os.makedirs(objdir, exist_ok=True)
    procs, objs = [], []
    for src in B.sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        procs.append(subprocess.Popen([B.NVCC] + B.FLAGS + defs + ["-c", src, "-o", obj]))
    assert all(p.wait() == 0 for p in procs)
    out = os.path.join(B.HERE, "libcdnet_b200_%s.so" % name)
    subprocess.check_call([B.NVCC, "-shared", "-o", out] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ldl"])
    print(out)


if __name__ == "__main__":
    main()
