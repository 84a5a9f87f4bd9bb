"""ctypes view of oracle/_build/liboracle.so (TEST INFRASTRUCTURE ONLY)."""
import ctypes

import numpy as np

from . import build as _build

_lib = None


def _get():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_build.build())
        for name in ("ws_flood_stable", "ws_flood_heap"):
            fn = getattr(_lib, name)
            fn.restype = ctypes.c_int
            fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                           ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        _lib.conv11_fma.restype = None
        _lib.conv11_fma.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                    ctypes.c_void_p, ctypes.c_void_p]
    return _lib


def ws_flood(image, markers, mask, order="stable"):
    """Marker-controlled flood (see ws_flood.c).  image any real dtype -> float64,
    markers -> int32, mask -> bool.  Returns int32 labels."""
    img = np.ascontiguousarray(image, dtype=np.float64)
    mk = np.ascontiguousarray(markers).astype(np.int32)
    ms = np.ascontiguousarray(np.asarray(mask) != 0).astype(np.uint8)
    assert img.ndim == 2 and img.shape == mk.shape == ms.shape
    out = np.empty(img.shape, dtype=np.int32)
    fn = getattr(_get(), "ws_flood_" + order)
    rc = fn(img.ctypes.data, mk.ctypes.data, ms.ctypes.data, img.shape[0], img.shape[1],
            out.ctypes.data)
    if rc != 0:
        raise MemoryError("ws_flood: allocation failed")
    return out


def conv11_fma(img, ker, sel=None):
    """f32 [H,W] * f32 [2,11,11] -> f32 [2,H,W]; sequential FMA chain per pixel (ws_flood.c)."""
    img = np.ascontiguousarray(img, dtype=np.float32)
    ker = np.ascontiguousarray(ker, dtype=np.float32).reshape(2, 11, 11)
    out = np.empty((2,) + img.shape, dtype=np.float32)
    selp = None
    if sel is not None:
        sel = np.ascontiguousarray(np.asarray(sel) != 0).astype(np.uint8)
        selp = sel.ctypes.data
    _get().conv11_fma(img.ctypes.data, img.shape[0], img.shape[1], ker.ctypes.data,
                      out.ctypes.data, selp)
    return out
