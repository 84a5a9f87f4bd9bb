"""SIMT emulator: runs the kernel sources of cdnet_b200/csrc on the host for the CPU test tier.

TEST INFRASTRUCTURE ONLY.  `emulated_api()` builds tests/simt/_build/libcdnet_b200_simt.so (g++; the same
.cu files, compiled against shims of the CUDA headers, see include/cuda_runtime.h and simt_runtime.cpp)
and yields cdnet_b200.api re-pointed at it for the duration of the `with` block: the emulated library's
"device pointers" are host pointers, so the host layer is given CPU tensors, no-op streams/events and
unpinned staging buffers.  Everything else -- argument marshalling, workspace sizing, status handling,
dtype conventions, every kernel -- is the product's own code.

The product never does any of this: cdnet_b200 has no CPU path and raises CdnetError without a GPU
(tests/test_cabi_symbols.py checks that the package does not reference tests/ or oracle/).
"""
import contextlib
import ctypes
import os

_lib = None


def load():
    global _lib
    if _lib is None:
        from . import build as _build
        from cdnet_b200 import _cabi
        # CDNET_SIMT_LIB: an alternative build of the emulated library (e.g. one compiled with -fsanitize=address,
        # run under LD_PRELOAD=libasan.so, to catch out-of-bounds accesses of the kernels)
        L = ctypes.CDLL(os.environ.get("CDNET_SIMT_LIB") or _build.build())
        for name, (res, args) in _cabi.SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class _NoStream(object):
    cuda_stream = None

    def __init__(self, *a, **k):
        pass

    def wait_stream(self, s):
        pass

    def wait_event(self, e):
        pass

    def synchronize(self):
        pass


class _NoEvent(object):
    def __init__(self, *a, **k):
        pass

    def record(self, s=None):
        pass

    def synchronize(self):
        pass


class _CudaProxy(object):
    Stream, Event = _NoStream, _NoEvent

    def __init__(self):
        self._cur = _NoStream()

    def is_available(self):
        return True

    def current_device(self):
        return 0

    def current_stream(self, device=None):
        return self._cur

    def stream(self, s):
        return contextlib.nullcontext()

    def synchronize(self, device=None):
        pass

    def set_device(self, d):
        pass

    def device(self, d):
        return contextlib.nullcontext()


class _TorchProxy(object):
    """`torch` as seen by cdnet_b200's host layer under emulation: no pinning, no CUDA streams"""

    def __init__(self, torch):
        self._t = torch
        self.cuda = _CudaProxy()

    def __getattr__(self, name):
        return getattr(self._t, name)

    def device(self, kind, index=None):
        return self._t.device("cpu")

    def empty(self, *a, **k):
        k.pop("pin_memory", None)
        return self._t.empty(*a, **k)

    def zeros(self, *a, **k):
        k.pop("pin_memory", None)
        return self._t.zeros(*a, **k)


@contextlib.contextmanager
def emulated_api():
    import torch
    from cdnet_b200 import _cabi, api, metrics, sharded, training
    L = load()
    cpu = torch.device("cpu")
    proxy = _TorchProxy(torch)
    mods = [api, metrics, training]
    saved = [(_cabi, "_lib", _cabi._lib), (api, "_device", api._device), (metrics, "_device", metrics._device),
             (training, "_device", training._device),
             (torch.Tensor, "record_stream", torch.Tensor.record_stream),
             (sharded.CudaBackend, "_st", sharded.CudaBackend._st)]
    saved += [(m, "torch", m.torch) for m in mods if hasattr(m, "torch")]
    plans = dict(getattr(api, "_plans", {}))
    try:
        _cabi._lib = L
        api._device = metrics._device = training._device = lambda device=None: cpu
        sharded.CudaBackend._st = lambda self: None
        torch.Tensor.record_stream = lambda self, s: None
        for m in mods:
            if hasattr(m, "torch"):
                m.torch = proxy
        api._plans.clear()
        api._ws_cache.clear()
        if os.environ.get("CDNET_SIMT_WS_EXACT"):
            # hand every call EXACTLY *_workspace_bytes(): a size formula that undercounts fails with E_WORKSPACE
            # instead of hiding in the 5 % slack the host layer normally adds
            saved.append((api, "_workspace", api._workspace))
            api._workspace = lambda nbytes, dev: torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=dev)
            for m in (metrics, training):
                saved.append((m, "_workspace", m._workspace))
                m._workspace = api._workspace
        fill = os.environ.get("CDNET_SIMT_WS_FILL")
        if fill:
            # initcheck: hand every call a workspace (and every torch.empty result) full of a junk byte, so a kernel
            # that reads scratch it has not written (fresh host pages are zero, recycled device memory is not)
            # changes its answer
            junk = int(fill, 0) & 0xff
            plain_ws = api._workspace
            saved.append((api, "_workspace", plain_ws))

            def dirty_ws(nbytes, dev):
                buf = plain_ws(nbytes, dev)
                buf.fill_(junk)
                return buf
            api._workspace = dirty_ws
            for m in (metrics, training):
                saved.append((m, "_workspace", m._workspace))
                m._workspace = dirty_ws
            real_empty_fn = proxy.empty

            def dirty_empty(*a, **k):
                t = real_empty_fn(*a, **k)
                t.view(torch.uint8).fill_(junk) if t.numel() else None
                return t
            proxy.empty = dirty_empty
            be_empty = sharded.CudaBackend.empty
            saved.append((sharded.CudaBackend, "empty", be_empty))

            def dirty_be_empty(self, shape, dtype):
                t = be_empty(self, shape, dtype)
                t.view(torch.uint8).fill_(junk) if t.numel() else None
                return t
            sharded.CudaBackend.empty = dirty_be_empty
        if os.environ.get("CDNET_SIMT_LIB"):
            # sanitizer builds may put guard gaps between the workspace slices: leave room for them
            real_ws = api._workspace
            saved.append((api, "_workspace", real_ws))
            api._workspace = lambda nbytes, dev: real_ws(nbytes + (1 << 16), dev)
            for m in (metrics, training):
                saved.append((m, "_workspace", m._workspace))
                m._workspace = api._workspace
            real_empty = sharded.CudaBackend.empty
            saved.append((sharded.CudaBackend, "empty", real_empty))
            sharded.CudaBackend.empty = lambda self, shape, dtype: (
                real_empty(self, (shape[0] + (1 << 16),), dtype) if dtype == "uint8" and len(shape) == 1
                else real_empty(self, shape, dtype))
        yield api
    finally:
        for obj, name, val in saved:
            setattr(obj, name, val)
        api._plans.clear()
        api._plans.update(plans)
        api._ws_cache.clear()
