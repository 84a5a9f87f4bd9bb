"""CPU: the SIMT emulator (tests/simt) tested on its own -- barriers, warp collectives with early exits and
partial masks, the schedule permutations that expose a missing barrier, deadlock detection, and the intrinsics
whose host restatement is not trivial (double -> half rounding, PRMT, funnel shifts)."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import HERE

SIMT = os.path.join(HERE, "simt")


@pytest.fixture(scope="module")
def lib_path():
    out = os.path.join(SIMT, "_build", "selftest")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libselftest.so")
    srcs = [os.path.join(SIMT, "selftest", "selftest.cu"), os.path.join(SIMT, "simt_runtime.cpp")]
    deps = srcs + [os.path.join(SIMT, "include", "cuda_runtime.h"), os.path.join(SIMT, "include", "cuda_fp16.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call([os.environ.get("CXX", "g++"), "-std=c++17", "-O1", "-g", "-fPIC", "-pthread", "-shared",
                               "-ffp-contract=off", "-fno-strict-aliasing", "-w", "-I", os.path.join(SIMT, "include"),
                               "-x", "c++"] + srcs + ["-o", so])
    return so


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def test_barriers_shuffles_atomics(lib_path):
    L = ctypes.CDLL(lib_path)
    x = np.random.default_rng(0).integers(-1000, 1000, size=100003).astype(np.int32)
    out = np.zeros(1, np.int32)
    L.st_block_sum(_ptr(x), _ptr(out), ctypes.c_int(x.size))
    assert int(out[0]) == int(x.sum())


def test_warp_collectives_with_exits_and_partial_masks(lib_path):
    L = ctypes.CDLL(lib_path)
    out = np.zeros(64 * 4, np.uint32)
    L.st_warp_ops(_ptr(out))
    out = out.reshape(64, 4)
    for t in range(40):
        lanes = range(32) if t < 32 else range(8)
        base = 0 if t < 32 else 32
        ballot = sum(1 << l for l in lanes if ((base + l) & 1) == 0)
        match = sum(1 << l for l in lanes if (base + l) % 3 == t % 3)
        rmax = max((base + l) * 7 % 13 for l in lanes)
        sub = sum(base + l for l in range(8)) if (t & 31) < 8 else 0
        assert list(out[t]) == [ballot, match, rmax, sub], t
    assert not out[40:].any()


def test_half_rounding_matches_numpy(lib_path):
    L = ctypes.CDLL(lib_path)
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.standard_normal(20000) * 10.0 ** rng.integers(-9, 6, 20000),
                        np.arange(0, 70000, 0.37), [0.0, -0.0, 65504.0, 65519.9, 65520.0, 1e9, -1e9, 5.96e-8, 2.98e-8,
                                                    2.9802322387695312e-08, 6.1e-5, np.inf, -np.inf]])
    # ties: exactly between two halfs
    h = rng.integers(0, 0x7bff, 5000).astype(np.uint16)
    mid = (h.view(np.float16).astype(np.float64) + (h + 1).astype(np.uint16).view(np.float16).astype(np.float64)) / 2
    x = np.ascontiguousarray(np.concatenate([x, mid, -mid]))
    out = np.zeros(x.size, np.uint16)
    L.st_half(_ptr(x), _ptr(out), ctypes.c_int(x.size))
    with np.errstate(over="ignore"):
        ref = x.astype(np.float16).view(np.uint16)
    assert np.array_equal(out, ref), np.flatnonzero(out != ref)[:5]


def test_bit_intrinsics(lib_path):
    L = ctypes.CDLL(lib_path)
    rng = np.random.default_rng(2)
    n = 4096
    a, b, s = (rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32) for _ in range(3))
    a[:4] = [0, 1, 0x80000000, 0xffffffff]
    b[:4] = [0, 0x10, 0x80000000, 1]
    out = np.zeros(4 * n, np.uint32)
    L.st_bits(_ptr(a), _ptr(b), _ptr(s), _ptr(out), ctypes.c_int(n))
    out = out.reshape(n, 4)
    for i in range(n):
        ai, bi, si = int(a[i]), int(b[i]), int(s[i])
        src = (bi << 32) | ai
        perm = 0
        for k in range(4):
            sel = (si >> (4 * k)) & 0xf
            byte = (src >> (8 * (sel & 7))) & 0xff
            if sel & 8:
                byte = 0xff if byte & 0x80 else 0
            perm |= byte << (8 * k)
        sh = si & 31
        fl = ((src << sh) >> 32) & 0xffffffff
        fr = (src >> sh) & 0xffffffff
        clz = 32 - ai.bit_length()
        ffs = (bi & -bi).bit_length()
        assert [int(v) for v in out[i]] == [perm, fl, fr, clz | (ffs << 8) | (bin(si).count("1") << 16)], i


_RACY = """
import ctypes, sys, numpy as np
L = ctypes.CDLL(sys.argv[1])
x = np.arange(1, 65, dtype=np.int32); out = np.zeros(64, np.int32)
L.st_racy_shift(x.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p))
print(",".join(str(int(v)) for v in out))
"""


def test_schedule_permutations_expose_a_missing_barrier(lib_path):
    """the schedule is fixed per process (SIMT_SCHEDULE): run the racy kernel in three child processes"""
    res = {}
    for mode in ("", "reverse", "random:5"):
        env = dict(os.environ, SIMT_SCHEDULE=mode, SIMT_WORKERS="1")
        res[mode] = subprocess.check_output([sys.executable, "-c", _RACY, lib_path], env=env).decode().strip()
    want = ",".join(str((t + 1) % 64 + 1) for t in range(64))
    assert len(set(res.values())) > 1, "a kernel with a missing barrier must not give one answer under every schedule"
    assert any(v != want for v in res.values())


def test_deadlock_is_reported(lib_path):
    code = "import ctypes, sys; ctypes.CDLL(sys.argv[1]).st_deadlock()"
    p = subprocess.run([sys.executable, "-c", code, lib_path], capture_output=True)
    assert p.returncode != 0 and b"simt: deadlock in k_deadlock" in p.stderr
