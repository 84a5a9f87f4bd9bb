"""Host-side mirror of the reference's callables for the geometry hot path.

Same names, argument meaning, return dtypes/shapes and error behaviour as honglianghe/CDNet, but
every computation runs in hand-written sm_100a kernels behind the C ABI of libcdnet_b200.so
(include/cdnet_b200.h).  PyTorch is used only for device memory, pinned staging and streams.

numpy in / numpy out ("drop-in") functions:
    generate_dd_map, circshift                data_prepare/getDirectionDiffMap.py:14-108
    process                                   postproc_other.py:15-54
    dam_postprocess                           test_dam.py:455-563 (inline block, wrapped)
    plain_postprocess                         test.py:270-295   (inline block, wrapped)
    label, binary_fill_holes, remove_small_objects, dilation, distance_transform_edt
                                              the scipy / scikit-image calls on the path
Device-resident batched variants (`*_cuda`, torch CUDA tensors [B,...]) keep the data on the GPU
between the CNN and the post-processing, and `DamPostprocessPlan` pre-allocates pinned staging +
device buffers for repeated host-buffer calls (what bench.py's e2e measures).
"""
import collections
import functools
import types

import numpy as np
import torch

from . import _cabi
from ._cabi import CdnetError, check

_ws_cache = collections.OrderedDict()  # (device index, stream handle) -> scratch buffer
_WS_CACHE_MAX = 8
_dev_checked = set()


def _device(device=None):
    if not torch.cuda.is_available():
        raise CdnetError("cdnet_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    if isinstance(device, torch.device):
        if device.type != "cuda":
            raise CdnetError("expected a CUDA device, got %s" % device)
        idx = device.index
    else:
        idx = device
    dev = torch.device("cuda", torch.cuda.current_device() if idx is None else int(idx))
    if dev.index not in _dev_checked:
        if not _cabi.lib().cdnet_device_ok(dev.index):
            raise CdnetError("device %s is not a compute-capability 10.x GPU; libcdnet_b200 is sm_100a-only" % dev)
        _dev_checked.add(dev.index)
    return dev


def _workspace(nbytes, dev):
    """scratch for one library call: one growing buffer per (device, stream), so that calls issued on different
    streams (or from different threads with their own streams) never share scratch memory"""
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        _ws_cache.pop(key, None)
        buf = None
        buf = torch.empty(int(nbytes * 1.05) + 4096, dtype=torch.uint8, device=dev)
        while len(_ws_cache) >= _WS_CACHE_MAX:
            _ws_cache.popitem(last=False)
    else:
        _ws_cache.pop(key)
    _ws_cache[key] = buf  # most recently used last
    return buf


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _first_cuda_tensor(args, kwargs):
    for x in list(args) + list(kwargs.values()):
        if isinstance(x, torch.Tensor):
            if x.is_cuda:
                return x
        elif isinstance(x, (list, tuple)) and x and isinstance(x[0], torch.Tensor) and x[0].is_cuda:
            return x[0]
    return None


def _on_tensor_device(fn):
    """Device guard: the library launches on the CURRENT device and on its current stream, so a call whose tensors
    live on another GPU runs inside `torch.cuda.device(tensor.device)` (stream, scratch memory and kernels then all
    belong to the tensors' device)."""
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        t = _first_cuda_tensor(args, kwargs)
        if t is None or t.device.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(t.device):
            return fn(*args, **kwargs)
    return wrapper


def guard_public_functions(ns):
    """wrap every public function (and static method of every public class) of a module namespace"""
    for name, obj in list(ns.items()):
        if name.startswith("_") or getattr(obj, "__module__", None) != ns.get("__name__"):
            continue
        if isinstance(obj, types.FunctionType):
            ns[name] = _on_tensor_device(obj)
        elif isinstance(obj, type):
            for k, v in list(vars(obj).items()):
                if isinstance(v, staticmethod) and not k.startswith("__"):
                    setattr(obj, k, staticmethod(_on_tensor_device(v.__func__)))


def _ptr(t):
    return None if t is None else t.data_ptr()


def _cu8(t):
    return t.contiguous() if t.dtype == torch.uint8 else t.to(torch.uint8).contiguous()


def launch_count():
    return int(_cabi.lib().cdnet_launch_count())


# =====================================================================================================
# device-resident batched API
# =====================================================================================================
def ddm_cuda(cls, direction_classes, return_status=False):
    """cls uint8 [B,H,W] CUDA -> float32 [B,H,W] (generate_dd_map per tile)."""
    L = _cabi.lib()
    dev = _device(cls.device)
    cls = _cu8(cls)
    B, H, W = cls.shape
    out = torch.empty((B, H, W), dtype=torch.float32, device=dev)
    status = torch.zeros((B,), dtype=torch.int32, device=dev)
    nb = L.cdnet_ddm_workspace_bytes(B, H, W)
    ws = _workspace(nb, dev)
    check(L.cdnet_ddm(_ptr(cls), _ptr(out), _ptr(status), B, H, W, int(direction_classes), _ptr(ws), ws.numel(),
                      _stream()), "cdnet_ddm")
    return (out, status) if return_status else out


def label_cuda(mask, connectivity=4, return_num=False):
    """mask [B,H,W] (non-zero = foreground) -> int32 labels, raster-first ids."""
    L = _cabi.lib()
    dev = _device(mask.device)
    m = _cu8(mask != 0) if mask.dtype != torch.uint8 else mask.contiguous()
    B, H, W = m.shape
    out = torch.empty((B, H, W), dtype=torch.int32, device=dev)
    n = torch.zeros((B,), dtype=torch.int32, device=dev)
    nb = L.cdnet_ccl_workspace_bytes(B, H, W)
    ws = _workspace(nb, dev)
    check(L.cdnet_ccl(_ptr(m), _ptr(out), _ptr(n), B, H, W, int(connectivity), _ptr(ws), ws.numel(), _stream()),
          "cdnet_ccl")
    return (out, n) if return_num else out


def label_values_cuda(ids, return_num=False):
    """ids uint8 [B,H,W] -> int32 labels of the 8-connected components of equal non-zero value, raster-first ids
    (skimage.measure.label of a multi-valued image)."""
    L = _cabi.lib()
    dev = _device(ids.device)
    assert ids.dtype == torch.uint8
    m = ids.contiguous()
    B, H, W = m.shape
    out = torch.empty((B, H, W), dtype=torch.int32, device=dev)
    n = torch.zeros((B,), dtype=torch.int32, device=dev)
    nb = L.cdnet_ccl_workspace_bytes(B, H, W)
    ws = _workspace(nb, dev)
    check(L.cdnet_label_values(_ptr(m), _ptr(out), _ptr(n), B, H, W, _ptr(ws), ws.numel(), _stream()),
          "cdnet_label_values")
    return (out, n) if return_num else out


def fill_holes_cuda(mask):
    L = _cabi.lib()
    dev = _device(mask.device)
    m = _cu8(mask != 0) if mask.dtype != torch.uint8 else mask.contiguous()
    B, H, W = m.shape
    out = torch.empty_like(m)
    nb = L.cdnet_fill_holes_workspace_bytes(B, H, W)
    ws = _workspace(nb, dev)
    check(L.cdnet_fill_holes(_ptr(m), _ptr(out), B, H, W, _ptr(ws), ws.numel(), _stream()), "cdnet_fill_holes")
    return out


def remove_small_mask_cuda(mask, min_size):
    L = _cabi.lib()
    dev = _device(mask.device)
    m = _cu8(mask != 0) if mask.dtype != torch.uint8 else mask.contiguous()
    B, H, W = m.shape
    out = torch.empty_like(m)
    nb = L.cdnet_remove_small_mask_workspace_bytes(B, H, W)
    ws = _workspace(nb, dev)
    check(L.cdnet_remove_small_mask(_ptr(m), _ptr(out), B, H, W, int(min_size), _ptr(ws), ws.numel(), _stream()),
          "cdnet_remove_small_mask")
    return out


def remove_small_labels_cuda(labels, min_size):
    """int32 [B,H,W] labels (values in [0, H*W]) -> copy with small labels zeroed."""
    L = _cabi.lib()
    dev = _device(labels.device)
    out = labels.to(torch.int32).contiguous().clone()
    B, H, W = out.shape
    nb = L.cdnet_remove_small_labels_workspace_bytes(B, H, W)
    ws = _workspace(nb, dev)
    check(L.cdnet_remove_small_labels(_ptr(out), B, H, W, int(min_size), _ptr(ws), ws.numel(), _stream()),
          "cdnet_remove_small_labels")
    return out


def label_dilate_cuda(labels, radius, out_dtype=torch.int32):
    L = _cabi.lib()
    dev = _device(labels.device)
    lab = labels.to(torch.int32).contiguous()
    B, H, W = lab.shape
    out = torch.empty((B, H, W), dtype=out_dtype, device=dev)
    check(L.cdnet_label_dilate(_ptr(lab), _ptr(out), out.element_size(), B, H, W, int(radius), _stream()),
          "cdnet_label_dilate")
    return out


def edt_cuda(mask, return_squared=False):
    """exact EDT: float64 [B,H,W] (and the int32 squared distances)."""
    L = _cabi.lib()
    dev = _device(mask.device)
    m = _cu8(mask != 0) if mask.dtype != torch.uint8 else mask.contiguous()
    B, H, W = m.shape
    d2 = torch.empty((B, H, W), dtype=torch.int32, device=dev)
    dist = torch.empty((B, H, W), dtype=torch.float64, device=dev)
    nb = L.cdnet_edt_workspace_bytes(B, H, W)
    ws = _workspace(nb, dev)
    check(L.cdnet_edt(_ptr(m), _ptr(d2), _ptr(dist), B, H, W, _ptr(ws), ws.numel(), _stream()), "cdnet_edt")
    return (dist, d2) if return_squared else dist


_NO_BACKGROUND_MSG = "list.remove(x): x not in list"  # what postproc_other.py:19 raises on a mask without background


def ws_contested_pixels(status):
    """status int32 [B] of a watershed-path call -> per-tile count of mask pixels whose label depends (to first order) on
    the order in which two age-0 markers of EQUAL priority are popped -- the one place where scikit-image's heap and
    this build's raster order can differ.  0 = the tile is independent of that order."""
    return (status >> _cabi.S_WS_CONTESTED_SHIFT) & 0xffffff


def process_cuda(pred01, min_size=10, ws=True, return_status=False):
    """pred01 uint8 [B,H,W] already binarised -> int32 labels (postproc_other.process).  With ws=True a tile
    without any background pixel sets CDNET_S_NO_BACKGROUND in status[b] (the reference raises ValueError)."""
    L = _cabi.lib()
    dev = _device(pred01.device)
    m = _cu8(pred01)
    B, H, W = m.shape
    out = torch.empty((B, H, W), dtype=torch.int32, device=dev)
    status = torch.zeros((B,), dtype=torch.int32, device=dev)
    nb = L.cdnet_ws_postproc_workspace_bytes(B, H, W)
    wsb = _workspace(nb, dev)
    check(L.cdnet_ws_postproc(_ptr(m), _ptr(out), _ptr(status), B, H, W, int(min_size), 1 if ws else 0, _ptr(wsb),
                              wsb.numel(), _stream()), "cdnet_ws_postproc")
    return (out, status) if return_status else out


def dam_postprocess_cuda(dcm, prob, point, direction_classes=9, min_area=20, radius=2, postproc=0,
                         write_prob=False, out_dtype=None, out=None, status=None):
    """dcm uint8 [B,8,H,W] (or [B,1,H,W]: single-map variant, test_dam.py:499-502), prob float32 [B,3,H,W],
    point float32 [B,1,H,W] (CUDA) -> labels [B,H,W].
    out_dtype defaults to what the reference returns: int64 (measure.label) for postproc 0, int32
    (process) for postproc 1.  Returns (labels, status)."""
    L = _cabi.lib()
    dev = _device(dcm.device)
    assert dcm.dtype == torch.uint8 and prob.dtype == torch.float32 and point.dtype == torch.float32
    assert dcm.is_contiguous() and prob.is_contiguous() and point.is_contiguous()
    B, T, H, W = dcm.shape
    assert T in (1, 8) and tuple(prob.shape) == (B, 3, H, W) and point.numel() == B * H * W
    if out_dtype is None:
        out_dtype = torch.int64 if int(postproc) == 0 else torch.int32
    if out is None:
        out = torch.empty((B, H, W), dtype=out_dtype, device=dev)
    if status is None:
        status = torch.empty((B,), dtype=torch.int32, device=dev)
    nb = L.cdnet_dam_postproc_workspace_bytes(B, H, W)
    ws = _workspace(nb, dev)
    check(L.cdnet_dam_postproc(_ptr(dcm), T, _ptr(prob), _ptr(point), _ptr(out), out.element_size(), _ptr(status), B,
                               H, W, int(direction_classes), int(min_area), int(radius), int(postproc),
                               1 if write_prob else 0, _ptr(ws), ws.numel(), _stream()), "cdnet_dam_postproc")
    return out, status


def plain_postprocess_cuda(prob, min_area=20, radius=2, postproc=0, multi_class=True, out_dtype=None):
    L = _cabi.lib()
    dev = _device(prob.device)
    assert prob.dtype == torch.float32 and prob.is_contiguous()
    B, C, H, W = prob.shape
    if out_dtype is None:
        out_dtype = torch.int64 if int(postproc) == 0 else torch.int32
    out = torch.empty((B, H, W), dtype=out_dtype, device=dev)
    status = torch.empty((B,), dtype=torch.int32, device=dev)
    nb = L.cdnet_plain_postproc_workspace_bytes(B, H, W)
    ws = _workspace(nb, dev)
    check(L.cdnet_plain_postproc(_ptr(prob), C, _ptr(out), out.element_size(), _ptr(status), B, H, W,
                                 1 if multi_class else 0, int(min_area), int(radius), int(postproc), _ptr(ws),
                                 ws.numel(), _stream()), "cdnet_plain_postproc")
    return out, status


# =====================================================================================================
# numpy drop-ins (reference signatures)
# =====================================================================================================
def _h2d(a, dtype=None):
    a = np.ascontiguousarray(a) if dtype is None else np.ascontiguousarray(a, dtype=dtype)
    return torch.from_numpy(a).to(_device(), non_blocking=False)


def generate_dd_map(label_direction, direction_classes):
    """data_prepare/getDirectionDiffMap.py:44-108.  [H,W] integer class map -> float32 [H,W];
    NaN for a constant map, like the reference's 0/0."""
    lab = np.asarray(label_direction)
    assert lab.ndim == 2
    # class ids outside 0..255 are unknown to label_to_vector (zero vector) -> any id >= n works
    cls = np.where((lab >= 0) & (lab < 255), lab, 255).astype(np.uint8) if lab.dtype != np.uint8 else lab
    out = ddm_cuda(_h2d(cls)[None], direction_classes)
    return out[0].cpu().numpy()


def circshift(matrix_ori, direction, shiftnum1, shiftnum2):
    """data_prepare/getDirectionDiffMap.py:14-42: zero-filled shift of a [C,H,W] array."""
    m = np.ascontiguousarray(matrix_ori)
    assert m.ndim == 3 and m.dtype.itemsize in (1, 2, 4, 8)
    L = _cabi.lib()
    view = {1: np.uint8, 2: np.uint16, 4: np.int32, 8: np.int64}[m.dtype.itemsize]
    src = _h2d(m.view(view))
    dst = torch.empty_like(src)
    C, H, W = m.shape
    check(L.cdnet_circshift(_ptr(src), _ptr(dst), C, H, W, m.dtype.itemsize, int(direction), int(shiftnum1),
                            int(shiftnum2), _stream()), "cdnet_circshift")
    return dst.cpu().numpy().view(m.dtype)


def label(input, connectivity=None, return_num=False):
    """skimage.measure.label (default: 8-connected, int64; components of equal value for a multi-valued uint8-range
    image); pass connectivity=1 for scipy.ndimage.label semantics on a binary image (4-connected, int32)."""
    x = np.asarray(input)
    conn = 8 if (connectivity is None or connectivity == 2) else 4
    if x.dtype != bool and x.size and x.max() > 1:
        # measure.label joins pixels of EQUAL value: an image with more than one distinct non-zero value takes the
        # value-aware kernel.  The decision is made on the device (256-bin presence table of cdnet_label_stats);
        # only images outside the uint8 range -- unsupported in multi-valued form -- are inspected on the host.
        if x.ndim == 2 and x.min() >= 0 and x.max() <= 255:
            d = _h2d(x.astype(np.uint8))[None]
            L = _cabi.lib()
            pres = torch.empty((1, 256), dtype=torch.int32, device=d.device)
            fg = torch.empty((1,), dtype=torch.int32, device=d.device)
            check(L.cdnet_label_stats(_ptr(d), _ptr(pres), _ptr(fg), 1, x.shape[0], x.shape[1], _stream()),
                  "cdnet_label_stats")
            multi = int((pres[0, 1:] != 0).sum()) > 1
            if multi and conn != 8:
                raise CdnetError("multi-valued label images are supported as 2-D, 8-connected, values 0..255")
            if multi:
                lab, n = label_values_cuda(d, return_num=True)
                res = lab[0].cpu().numpy().astype(np.int64)
                return (res, int(n[0])) if return_num else res
        elif np.unique(x[x != 0]).size > 1:
            raise CdnetError("multi-valued label images are supported as 2-D, 8-connected, values 0..255")
    lab, n = label_cuda(_h2d((x != 0).astype(np.uint8))[None], conn, return_num=True)
    res = lab[0].cpu().numpy()
    res = res.astype(np.int64) if conn == 8 else res
    return (res, int(n[0])) if return_num else res


def binary_fill_holes(input):
    """scipy.ndimage.binary_fill_holes -> bool [H,W]."""
    x = np.asarray(input)
    return fill_holes_cuda(_h2d((x != 0).astype(np.uint8))[None])[0].cpu().numpy().astype(bool)


def remove_small_objects(ar, min_size=64, connectivity=1):
    """skimage.morphology.remove_small_objects (bool: 4-connected components; int: values are labels)."""
    ar = np.asarray(ar)
    if min_size == 0:
        return ar.copy()
    if ar.dtype == bool:
        assert connectivity == 1
        return remove_small_mask_cuda(_h2d(ar.astype(np.uint8))[None], min_size)[0].cpu().numpy().astype(bool)
    if not np.issubdtype(ar.dtype, np.integer):
        raise TypeError("Only bool or integer image types are supported. Got %s." % ar.dtype)
    if ar.size and ar.min() < 0:
        raise ValueError("Negative value labels are not supported.")
    if ar.size and ar.max() > ar.size:
        raise CdnetError("label values above H*W are not supported by the CUDA histogram")
    out = remove_small_labels_cuda(_h2d(ar.astype(np.int32))[None], min_size)[0].cpu().numpy()
    return out.astype(ar.dtype)


def dilation(image, selem=None, radius=None):
    """skimage.morphology.dilation(label_image, disk(radius)); selem=None -> the cross (radius 1)."""
    img = np.asarray(image)
    if radius is None:
        radius = 1 if selem is None else (np.asarray(selem).shape[0] - 1) // 2
        if selem is not None:
            r = radius
            yy, xx = np.mgrid[-r:r + 1, -r:r + 1]
            if not np.array_equal(np.asarray(selem) != 0, (xx * xx + yy * yy) <= r * r):
                raise CdnetError("only disk(r) footprints are supported")
    if img.size and img.min() < 0:
        raise CdnetError("label images must be non-negative")
    out = label_dilate_cuda(_h2d(img.astype(np.int32))[None], radius)[0].cpu().numpy()
    return out.astype(img.dtype)


def distance_transform_edt(input):
    """scipy.ndimage.distance_transform_edt -> float64 [H,W]."""
    x = np.asarray(input)
    return edt_cuda(_h2d((x != 0).astype(np.uint8))[None])[0].cpu().numpy()


def process(pred, model_mode, min_size=10, ws=True):
    """postproc_other.process (postproc_other.py:15-54).  Binarises `pred` IN PLACE like the reference
    (:31-32) and returns int32 labels (ids keep gaps)."""
    if model_mode == "dcan":
        raise NotImplementedError("the dcan branch (postproc_other.py:69-97) is out of scope")
    if model_mode == "micronet":
        raise NotImplementedError("the micronet tail (postproc_other.py:56-68) is out of scope")
    assert len(pred.shape) == 2, "Prediction shape is not HW"
    hi = pred > 0.5
    pred[hi] = 1
    pred[~hi] = 0
    if model_mode == "unet":
        ws = False
    if ws and hi.all():
        raise ValueError(_NO_BACKGROUND_MSG)  # gen_inst_dst_map: `nuc_list.remove(0)` (postproc_other.py:18-19)
    return process_cuda(_h2d(hi.astype(np.uint8))[None], min_size, ws)[0].cpu().numpy()


class DamPostprocessPlan(object):
    """Pre-allocated pinned staging + device buffers for repeated host-buffer calls of the
    direction-aware post-processing on B tiles of H x W (test_dam.py:455-563).

    Fill `h_dcm` [B,8,H,W] u8, `h_prob` [B,3,H,W] f32, `h_point` [B,1,H,W] f32 (pinned numpy views),
    call `run()`, read `h_labels` [B,H,W]."""

    def __init__(self, B, H, W, direction_classes=9, min_area=20, radius=2, postproc=0, write_prob=False,
                 device=None, out_dtype=None):
        self.dev = _device(device)
        self.args = (int(direction_classes), int(min_area), int(radius), int(postproc), bool(write_prob))
        if out_dtype is None:
            out_dtype = torch.int64 if int(postproc) == 0 else torch.int32
        pin = dict(pin_memory=True)
        self.t_dcm = torch.empty((B, 8, H, W), dtype=torch.uint8, **pin)
        self.t_prob = torch.empty((B, 3, H, W), dtype=torch.float32, **pin)
        self.t_point = torch.empty((B, 1, H, W), dtype=torch.float32, **pin)
        self.t_labels = torch.empty((B, H, W), dtype=out_dtype, **pin)
        self.t_status = torch.zeros((B,), dtype=torch.int32, **pin)
        self.h_dcm, self.h_prob, self.h_point = self.t_dcm.numpy(), self.t_prob.numpy(), self.t_point.numpy()
        self.h_labels, self.h_status = self.t_labels.numpy(), self.t_status.numpy()
        self.d_dcm = torch.empty_like(self.t_dcm, device=self.dev)
        self.d_prob = torch.empty_like(self.t_prob, device=self.dev)
        self.d_point = torch.empty_like(self.t_point, device=self.dev)
        self.d_labels = torch.empty_like(self.t_labels, device=self.dev)
        self.d_status = torch.zeros((B,), dtype=torch.int32, device=self.dev)
        self.h2d_bytes = self.t_dcm.numel() + 4 * self.t_prob.numel() + 4 * self.t_point.numel()
        self.d2h_bytes = self.t_labels.numel() * self.t_labels.element_size() + 4 * B
        # the boosted boundary channel (test_dam.py:536 updates prob_maps[2] in place) comes back in its OWN pinned
        # buffer: the staged inputs stay what the caller wrote, so a plan can be re-launched on the same inputs
        self.t_prob2 = torch.empty((B, H, W), dtype=torch.float32, **pin) if write_prob else None
        self.h_prob2 = self.t_prob2.numpy() if write_prob else None
        if write_prob:
            self.d2h_bytes += 4 * self.t_prob2.numel()
        self._s_in = self._s_out = None
        self._graph = None

    def launch(self, chunk=2):
        """H2D, kernels, D2H.  The batch is cut into chunks of `chunk` tiles: the pinned-host -> device copy
        of chunk i+1 (copy-in stream) and the device -> host copy of chunk i-1 (copy-out stream) overlap the
        kernels of chunk i (current stream).  On return the current stream has been made to wait for the last
        copy-out, so synchronising it (or recording an event on it) covers the whole step."""
        with torch.cuda.device(self.dev):
            self._launch(chunk)

    def _launch(self, chunk):
        B = self.t_dcm.shape[0]
        cur = torch.cuda.current_stream()
        if self._s_in is None:
            self._s_in, self._s_out = torch.cuda.Stream(self.dev), torch.cuda.Stream(self.dev)
        s_in, s_out = self._s_in, self._s_out
        s_in.wait_stream(cur)   # the previous step may still be reading the device inputs
        s_out.wait_stream(cur)
        dc, ma, ra, pp, wp = self.args
        for a in range(0, B, chunk):
            b = min(B, a + chunk)
            with torch.cuda.stream(s_in):
                self.d_dcm[a:b].copy_(self.t_dcm[a:b], non_blocking=True)
                self.d_prob[a:b].copy_(self.t_prob[a:b], non_blocking=True)
                self.d_point[a:b].copy_(self.t_point[a:b], non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record(s_in)
            cur.wait_event(ev_in)
            dam_postprocess_cuda(self.d_dcm[a:b], self.d_prob[a:b], self.d_point[a:b], dc, ma, ra, pp, wp,
                                 out=self.d_labels[a:b], status=self.d_status[a:b])
            ev_k = torch.cuda.Event()
            ev_k.record(cur)
            s_out.wait_event(ev_k)
            with torch.cuda.stream(s_out):
                self.t_labels[a:b].copy_(self.d_labels[a:b], non_blocking=True)
                self.t_status[a:b].copy_(self.d_status[a:b], non_blocking=True)
                if wp:
                    self.t_prob2[a:b].copy_(self.d_prob[a:b, 2], non_blocking=True)
        cur.wait_stream(s_out)

    def launch_device(self):
        """kernels only, inputs already resident in d_dcm / d_prob / d_point (write_prob then updates d_prob[:, 2]
        in place, as dam_postprocess_cuda does)."""
        dc, ma, ra, pp, wp = self.args
        with torch.cuda.device(self.dev):
            dam_postprocess_cuda(self.d_dcm, self.d_prob, self.d_point, dc, ma, ra, pp, wp, out=self.d_labels,
                                 status=self.d_status)

    def capture_graph(self, chunk=None):
        """Captures one launch() (H2D copies, every kernel, D2H copies) in a CUDA graph: replay() then costs one
        graph launch instead of ~20 stream operations -- what matters for the reference's call pattern of ONE tile
        per call.  Call run() once before (scratch memory and kernel attributes are set up outside the capture)."""
        B = self.t_dcm.shape[0]
        with torch.cuda.device(self.dev):
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._launch(B if chunk is None else chunk)
            self._graph = g
        return g

    def replay(self):
        self._graph.replay()

    def run(self):
        self.launch()
        torch.cuda.current_stream(self.dev).synchronize()
        if (self.h_status & _cabi.S_DDM_CONSTANT).any():
            # the reference: NaN direction-difference map -> `assert(np.min(enhanced_boundary) >= 0)` fails
            raise AssertionError("constant direction map: generate_dd_map is NaN (test_dam.py:535)")
        if (self.h_status & _cabi.S_NO_BACKGROUND).any():
            raise ValueError(_NO_BACKGROUND_MSG)  # process() on an all-inside tile (postproc_other.py:18-19)
        return self.h_labels


_plans = collections.OrderedDict()  # a few recently used single-tile plans (pinned staging is expensive to allocate)
_PLANS_MAX = 4


def _process_mode(postproc, model_name):
    """postproc 1 hands the mask to postproc_other.process(model_mode=model_name): 'unet' runs it without the
    watershed (postproc_other.py:35) = mode 2 of the C ABI; the micronet / dcan tails are out of scope"""
    postproc = int(postproc)
    if postproc == 1:
        if model_name in ("micronet", "dcan"):
            raise NotImplementedError("postproc=1 with model_mode %r is out of scope" % model_name)
        if model_name == "unet":
            return 2
    return postproc


def dam_postprocess(prob_maps, point_maps, dcm_tta, direction_classes=9, min_area=20, radius=2, postproc=0,
                    model_name="modelName", mutate_prob=True, voting_first=False):
    """test_dam.py:455-563 as a function.

    prob_maps  float32 [3,H,W]  (channel 2 is overwritten in place with the boosted boundary
                                probability, like the reference does at :536, unless mutate_prob=False)
    point_maps float32 [1,H,W]
    dcm_tta    uint8 [8,H,W] or [H,W,8]: prob_dcm, _hf, _vf, _hvf, _r90, _r90_hf, _r90_vf, _r90_hvf
    voting_first: the block's `voting_firt` switch (:471), off in the reference as shipped
    Returns pred_labeled [H,W]: int64 (postproc 0, measure.label) or int32 (postproc 1, process)."""
    postproc = _process_mode(postproc, model_name)
    prob = np.asarray(prob_maps)
    H, W = prob.shape[1:]
    dcm = np.asarray(dcm_tta)
    if dcm.shape == (H, W, 8) and dcm.shape != (8, H, W):
        dcm = np.moveaxis(dcm, 2, 0)
    if voting_first:
        # the reference's `voting_firt = 1` switch (test_dam.py:471-477): DcmVoting2 over the 8 maps, then ONE
        # direction-difference map instead of the mean of eight -- all on the device
        dev = _device()
        d8 = torch.from_numpy(np.ascontiguousarray(dcm, dtype=np.uint8)).to(dev)[None]
        p = torch.from_numpy(np.ascontiguousarray(prob, dtype=np.float32)).to(dev)[None]
        q = torch.from_numpy(np.ascontiguousarray(np.asarray(point_maps, dtype=np.float32).reshape(1, 1, H, W))).to(dev)
        labels, status = dam_postprocess_cuda(dcm_voting2_cuda(d8)[:, None].contiguous(), p, q, direction_classes,
                                              min_area, radius, postproc, write_prob=mutate_prob)
        st = int(status[0])
        if st & _cabi.S_DDM_CONSTANT:
            raise AssertionError("constant direction map: generate_dd_map is NaN (test_dam.py:535)")
        if st & _cabi.S_NO_BACKGROUND:
            raise ValueError(_NO_BACKGROUND_MSG)
        if mutate_prob and isinstance(prob_maps, np.ndarray):
            prob_maps[2, :, :] = p[0, 2].cpu().numpy()
        return labels[0].cpu().numpy()
    key = (H, W, int(direction_classes), int(min_area), int(radius), int(postproc), bool(mutate_prob),
           torch.cuda.current_device())
    plan = _plans.pop(key, None)
    if plan is None:
        while len(_plans) >= _PLANS_MAX:
            _plans.popitem(last=False)  # least recently used tile shape goes
        plan = DamPostprocessPlan(1, H, W, direction_classes, min_area, radius, postproc, write_prob=mutate_prob)
    _plans[key] = plan
    plan.h_dcm[0] = dcm
    plan.h_prob[0] = prob
    plan.h_point[0] = np.asarray(point_maps).reshape(1, H, W)
    labels = plan.run()[0].copy()
    if mutate_prob and isinstance(prob_maps, np.ndarray):
        prob_maps[2, :, :] = plan.h_prob2[0]
    return labels


def plain_postprocess(prob_maps, min_area=20, radius=2, postproc=0, model_name="modelName", multi_class=True):
    """test.py:270-295 as a function -> pred_labeled [H,W] (int64 for postproc 0, int32 for 1)."""
    postproc = _process_mode(postproc, model_name)
    prob = _h2d(np.asarray(prob_maps), np.float32)[None]
    out, status = plain_postprocess_cuda(prob, min_area, radius, postproc, multi_class)
    if int(status[0]) & _cabi.S_NO_BACKGROUND:
        raise ValueError(_NO_BACKGROUND_MSG)  # process() on an all-inside tile (postproc_other.py:18-19)
    return out[0].cpu().numpy()


def dcm_voting2_cuda(dcm):
    """dcm uint8 [B,8,H,W] -> voted direction class uint8 [B,H,W] (utils.py:1150-1159)."""
    L = _cabi.lib()
    dev = _device(dcm.device)
    d = _cu8(dcm)
    B, T, H, W = d.shape
    assert T == 8
    out = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
    check(L.cdnet_dcm_voting2(_ptr(d), _ptr(out), B, H, W, _stream()), "cdnet_dcm_voting2")
    return out


def DcmVoting2(direct_map):
    """utils.py:1150-1159: direct_map uint8 [H,W,8] -> int64 [H,W] (np.argmax of the per-class votes)."""
    dm = np.asarray(direct_map)
    assert dm.ndim == 3 and dm.shape[2] == 8
    d = _h2d(np.ascontiguousarray(np.moveaxis(dm, 2, 0)).astype(np.uint8))[None]
    return dcm_voting2_cuda(d)[0].cpu().numpy().astype(np.int64)


def direction_argmax_cuda(mask_logits, direction_logits):
    """Device-resident hand-off from the CNN (test_dam.py:984-1013): softmax of the
    3-class mask head and of the direction head, direction[0] *= mask[0], argmax -> (prob float32 [B,3,H,W],
    direction class uint8 [B,H,W]).  Stock torch ops on the tensors the CNN already holds on the GPU (this is
    exactly what the reference runs before its .cpu().numpy()); the point is that nothing leaves the device
    before dam_postprocess_cuda."""
    prob = torch.softmax(mask_logits, dim=1)
    dprob = torch.softmax(direction_logits, dim=1)
    dprob[:, 0] = dprob[:, 0] * prob[:, 0]
    return prob.contiguous(), torch.argmax(dprob, dim=1).to(torch.uint8).contiguous()


TTA_VARIANTS = ("id", "hf", "vf", "hvf", "r90", "r90_hf", "r90_vf", "r90_hvf")


def tta_merge_cuda(mask_logits, point, dir_logits):
    """Device-resident hand-off for test-time augmentation (test_dam.py:299-450 and get_probmaps :983-1013) in ONE
    kernel.  Each argument is a sequence of 8 CUDA float32 tensors, the raw network outputs of the variants in the
    order TTA_VARIANTS, in the variant's own frame: [B,3,h,w], [B,1,h,w], [B,C,h,w] (h,w = H,W for the flips,
    W,H for the rotated ones) -- or sequences of ONE tensor each when test-time augmentation is off.  Returns
    (prob float32 [B,3,H,W], point float32 [B,1,H,W], dcm uint8 [B,n,H,W]) in the original frame: exactly the
    inputs of dam_postprocess_cuda."""
    import ctypes
    L = _cabi.lib()
    nv = len(mask_logits)
    assert nv in (1, 8) and len(point) == nv and len(dir_logits) == nv
    dev = _device(mask_logits[0].device)
    B, three, H, W = mask_logits[0].shape
    C = dir_logits[0].shape[1]
    assert three == 3
    keep = []
    arrays = []
    for seq, ch in ((mask_logits, 3), (point, 1), (dir_logits, C)):
        ptrs = (ctypes.c_void_p * nv)()
        for v, t in enumerate(seq):
            shape = (B, ch, H, W) if v < 4 else (B, ch, W, H)
            if t.dim() == 3 and ch == 1:
                t = t[:, None]
            assert t.dtype == torch.float32 and tuple(t.shape) == shape, (TTA_VARIANTS[v], tuple(t.shape), shape)
            t = t.contiguous()
            keep.append(t)
            ptrs[v] = t.data_ptr()
        arrays.append(ptrs)
    prob = torch.empty((B, 3, H, W), dtype=torch.float32, device=dev)
    pt = torch.empty((B, 1, H, W), dtype=torch.float32, device=dev)
    dcm = torch.empty((B, nv, H, W), dtype=torch.uint8, device=dev)
    check(L.cdnet_tta_merge(arrays[0], arrays[1], arrays[2], nv, _ptr(prob), _ptr(pt), _ptr(dcm), B, H, W, int(C),
                            _stream()), "cdnet_tta_merge")
    return prob, pt, dcm


# =====================================================================================================
# target transform (my_transforms_direction.py:651-885)
# =====================================================================================================
def _gauss_weights():
    # scipy.ndimage._filters._gaussian_kernel1d(sigma=2, order=0, radius=int(4*2+0.5)) evaluated with the
    # host's numpy exactly as scipy does, so the weights are bit-identical to the reference's on this host
    x = np.arange(-8, 9)
    phi = np.exp(-0.5 / 4.0 * x ** 2)
    return np.ascontiguousarray(phi / phi.sum(), dtype=np.float64)


def label_stats_cuda(ids):
    """ids uint8 (or int32) [B,H,W] -> (n_distinct int32 [B], fg_count int32 [B]) on the device; for int32 ids n_distinct
    is capped at 3 (the reference only asks `> 2`)."""
    L = _cabi.lib()
    dev = _device(ids.device)
    if ids.dtype == torch.int32:
        ids = ids.contiguous()
        B, H, W = ids.shape
        nd = torch.empty((B,), dtype=torch.int32, device=dev)
        fg = torch.empty((B,), dtype=torch.int32, device=dev)
        scratch = torch.empty((B, 5), dtype=torch.int32, device=dev)
        check(L.cdnet_label_stats_i32(_ptr(ids), _ptr(nd), _ptr(fg), _ptr(scratch), B, H, W, _stream()), "cdnet_label_stats_i32")
        return nd, fg
    ids = _cu8(ids)
    B, H, W = ids.shape
    pres = torch.empty((B, 256), dtype=torch.int32, device=dev)
    fg = torch.empty((B,), dtype=torch.int32, device=dev)
    check(L.cdnet_label_stats(_ptr(ids), _ptr(pres), _ptr(fg), B, H, W, _stream()), "cdnet_label_stats")
    return pres.sum(dim=1), fg


def encode_targets_cuda(ids, instance_level=True, num_classes=8, want_inst=False, want_dir=False):
    """ids uint8 [B,H,W] (channel 0 of the label image) ->
    (ternary uint8 [B,H,W], point float16 [B,H,W], direction int64 [B,H,W][, inst int32][, dir f32 [B,H,W,2]]).
    instance_level: True / 1 = instance ids, False / 0 = {0,255} label (both out_c == 3); 2 / 3 = the same two input
    kinds for out_c != 3 (my_transforms_direction.py:721-739; for 3 pass max(channel 0, channel 1)); 4 / 5 =
    my_transforms.LabelEncoding with do_direction = 1.  int32 ids (labels loaded without the uint8 truncation) take the
    out_c == 3 forms (0 / 1) through cdnet_encode_targets_i32."""
    L = _cabi.lib()
    dev = _device(ids.device)
    wide = ids.dtype == torch.int32
    ids = ids.contiguous() if wide else _cu8(ids)
    if wide and int(instance_level) > 1:
        raise CdnetError("int32 label ids are supported for the out_c == 3 transform only")
    B, H, W = ids.shape
    ternary = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
    point = torch.empty((B, H, W), dtype=torch.float16, device=dev)
    direction = torch.empty((B, H, W), dtype=torch.int64, device=dev)
    inst = torch.empty((B, H, W), dtype=torch.int32, device=dev) if want_inst else None
    dirm = torch.empty((B, H, W, 2), dtype=torch.float32, device=dev) if want_dir else None
    status = torch.empty((B,), dtype=torch.int32, device=dev)
    nb = L.cdnet_encode_targets_workspace_bytes(B, H, W)
    ws = _workspace(nb, dev)
    gw = _gauss_weights()
    fn = L.cdnet_encode_targets_i32 if wide else L.cdnet_encode_targets
    check(fn(_ptr(ids), int(instance_level), _ptr(ternary), _ptr(point), _ptr(direction), _ptr(inst), _ptr(dirm),
             _ptr(status), B, H, W, int(num_classes), gw.ctypes.data, _ptr(ws), ws.numel(), _stream()),
          "cdnet_encode_targets")
    res = [ternary, point, direction]
    if want_inst:
        res.append(inst)
    if want_dir:
        res.append(dirm)
    return tuple(res)


def center_points_cuda(labels, max_label):
    """labels int32 [B,H,W] -> centres int32 [B, max_label+1, 2] (row, col); (-1,-1) for absent ids."""
    L = _cabi.lib()
    dev = _device(labels.device)
    lab = labels.to(torch.int32).contiguous()
    B, H, W = lab.shape
    out = torch.empty((B, int(max_label) + 1, 2), dtype=torch.int32, device=dev)
    nb = L.cdnet_center_points_workspace_bytes(B, H, W, int(max_label))
    ws = _workspace(nb, dev)
    check(L.cdnet_center_points(_ptr(lab), _ptr(out), B, H, W, int(max_label), _ptr(ws), ws.numel(), _stream()),
          "cdnet_center_points")
    return out


def get_centerpoint2(mask, n=None, m=None):
    """my_transforms_direction.py:651-685: [row, col] of the first raster-order pixel of maximum centerness
    of the (single) nucleus `mask > 0`; [-1, -1] for an empty mask."""
    mk = np.asarray(mask)
    if n is not None and m is not None:
        mk = mk[:n, :m]
    lab = _h2d((mk > 0).astype(np.int32))[None]
    c = center_points_cuda(lab, 1)[0, 1].cpu().numpy()
    return [int(c[0]), int(c[1])]


class LabelEncoding(object):
    """Drop-in for my_transforms_direction.LabelEncoding (:687-885): `LabelEncoding(out_c, radius,
    do_direction)(imgs)` with imgs = (img, weight_map, label) returns
    (img, weight_map, PIL 'L' ternary label {0,127,255}[, float16 point map, int64 direction classes]);
    out_c != 3 (options.py:42: multi_class off) gives a {0,255} label image without a boundary class.

    The number of direction classes is the reference's env `dt_num_classes` (default 8;
    data_prepare/SegFix_offset_helper.py:37-39) unless `num_classes` is given.  CUDA cannot be
    initialised in a forked DataLoader worker: apply this transform in the main process (or use
    workers started with 'spawn'); `encode_batch` takes a whole batch of label images at once."""

    def __init__(self, out_c=3, radius=1, do_direction=0, num_classes=None):
        import os
        self.out_c = out_c
        self.radius = 1  # the reference ignores its argument (:694)
        self.do_direction = do_direction
        self.num_classes = int(os.environ.get("dt_num_classes", 8)) if num_classes is None else int(num_classes)

    @staticmethod
    def _channel0(label):
        """channel 0 of the label image as the kernels take it: uint8 (what data_folder.py:26-37 delivers), or int32
        when the ids do not fit a byte (label images loaded with compat.data_folder.img_loader(keep_ids=True))"""
        if not isinstance(label, np.ndarray):
            label = np.array(label)
        inside = label if label.ndim == 2 else label[:, :, 0]
        if inside.dtype != np.uint8:
            if inside.size and inside.min() < 0:
                raise CdnetError("label ids must be non-negative")
            if inside.size and inside.max() > 255:
                if inside.max() > 2 ** 31 - 1:
                    raise CdnetError("label ids must fit int32")
                return np.ascontiguousarray(inside, dtype=np.int32)
            inside = inside.astype(np.uint8)
        return np.ascontiguousarray(inside)

    def encode_batch(self, labels):
        """labels: list of label images (same H x W) -> list of (ternary u8, point f16, direction int64)."""
        try:
            dev = _device()
        except RuntimeError as e:  # pragma: no cover
            raise CdnetError("LabelEncoding needs CUDA in this process (forked DataLoader workers cannot "
                             "initialise it; run the transform post-collate in the main process): %s" % e)
        ids = np.stack([self._channel0(l) for l in labels])
        d_ids = torch.from_numpy(ids).to(dev)
        ndist, _ = label_stats_cuda(d_ids)
        level = (ndist > 2).cpu().numpy()
        d_bin = d_ids
        if self.out_c != 3:
            # :721-739 index label[:, :, 0] (and [:, :, 1] for a {0,255} label): 2-D label images cannot be indexed
            for l in labels:
                if np.ndim(l) == 2:
                    raise IndexError("too many indices for array: array is 2-dimensional, but 3 were indexed")
            if not level.all():
                # new_label = 2 where channel 0 OR channel 1 exceeds 127.5 (:730-731)  <=>  max(ch0, ch1) > 127.5
                ch1 = np.stack([np.ascontiguousarray(np.asarray(l)[:, :, 1]).astype(np.uint8) for l in labels])
                d_bin = torch.maximum(d_ids, torch.from_numpy(ch1).to(dev))
        out = [None] * len(labels)
        for lv in (True, False):
            sel = np.nonzero(level == lv)[0]
            if sel.size == 0:
                continue
            src = d_ids if (lv or self.out_c == 3) else d_bin
            sub = src[torch.from_numpy(sel).to(dev)] if sel.size != len(labels) else src
            mode = (1 if lv else 0) if self.out_c == 3 else (2 if lv else 3)
            tern, point, direction, inst = encode_targets_cuda(sub, instance_level=mode, num_classes=self.num_classes,
                                                               want_inst=True)
            # the reference skips the first entry of np.unique(label_instance) as "background" (:797-800); on a
            # tile whose dilated instances leave no background pixel that drops a nucleus and its
            # `assert int(label_point.sum() / 255) == markers_len` (:836) fires
            # (for out_c != 3 the same slip only drops the smallest instance: the kernels reproduce that)
            if self.out_c == 3 and int(inst.reshape(inst.shape[0], -1).min(dim=1).values.max()) > 0:
                raise AssertionError("label_instance has no background pixel (my_transforms_direction.py:836)")
            tern, point, direction = tern.cpu().numpy(), point.cpu().numpy(), direction.cpu().numpy()
            for j, i in enumerate(sel):
                out[i] = (tern[j], point[j], direction[j])
        return out

    def __call__(self, imgs):
        from PIL import Image
        out_imgs = list(imgs)
        tern, point, direction = self.encode_batch([imgs[2]])[0]
        out_imgs[2] = Image.fromarray(tern)
        if self.do_direction == 1:
            out_imgs.append(point)
            out_imgs.append(direction)
        return tuple(out_imgs)


class EncodeTargetsPlan(object):
    """Pre-allocated pinned staging + device buffers for repeated host-buffer calls of the target transform on
    B label tiles of H x W (LabelEncoding path, my_transforms_direction.py:697-885, instance-level labels).

    Fill `h_ids` [B,H,W] uint8 (channel 0 of the label images), call `run()`, read `h_ternary` (uint8),
    `h_point` (float16) and `h_direction` (int64).  Copies of chunk i+1 / i-1 overlap the kernels of chunk i."""

    def __init__(self, B, H, W, num_classes=8, instance_level=True, device=None):
        self.dev = _device(device)
        self.num_classes, self.instance_level = int(num_classes), bool(instance_level)
        pin = dict(pin_memory=True)
        self.t_ids = torch.empty((B, H, W), dtype=torch.uint8, **pin)
        self.t_ternary = torch.empty((B, H, W), dtype=torch.uint8, **pin)
        self.t_point = torch.empty((B, H, W), dtype=torch.float16, **pin)
        self.t_direction = torch.empty((B, H, W), dtype=torch.int64, **pin)
        self.h_ids, self.h_ternary = self.t_ids.numpy(), self.t_ternary.numpy()
        self.h_point, self.h_direction = self.t_point.numpy(), self.t_direction.numpy()
        self.d_ids = torch.empty_like(self.t_ids, device=self.dev)
        self.h2d_bytes = self.t_ids.numel()
        self.d2h_bytes = self.t_ids.numel() * (1 + 2 + 8)
        self._s_in = self._s_out = None

    def launch(self, chunk=32):
        with torch.cuda.device(self.dev):
            self._launch(chunk)

    def _launch(self, chunk):
        B = self.t_ids.shape[0]
        cur = torch.cuda.current_stream()
        if self._s_in is None:
            self._s_in, self._s_out = torch.cuda.Stream(self.dev), torch.cuda.Stream(self.dev)
        s_in, s_out = self._s_in, self._s_out
        s_in.wait_stream(cur)
        s_out.wait_stream(cur)
        for a in range(0, B, chunk):
            b = min(B, a + chunk)
            with torch.cuda.stream(s_in):
                self.d_ids[a:b].copy_(self.t_ids[a:b], non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record(s_in)
            cur.wait_event(ev_in)
            tern, point, direction = encode_targets_cuda(self.d_ids[a:b], self.instance_level, self.num_classes)
            ev_k = torch.cuda.Event()
            ev_k.record(cur)
            s_out.wait_event(ev_k)
            with torch.cuda.stream(s_out):
                self.t_ternary[a:b].copy_(tern, non_blocking=True)
                self.t_point[a:b].copy_(point, non_blocking=True)
                self.t_direction[a:b].copy_(direction, non_blocking=True)
                # keep the device results alive until the copies have run
                for t in (tern, point, direction):
                    t.record_stream(s_out)
        cur.wait_stream(s_out)

    def run(self):
        self.launch()
        torch.cuda.current_stream(self.dev).synchronize()
        return self.h_ternary, self.h_point, self.h_direction


guard_public_functions(globals())
