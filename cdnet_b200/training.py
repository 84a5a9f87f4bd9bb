"""Training-side drop-ins (SURVEY.md section 8a rows A3/A4 as stand-alone operators, section 8f row 4).

    DTOffsetConfig, DTOffsetHelper      data_prepare/SegFix_offset_helper.py:21-46, 246-261, 286-341, 423-506
    Sobel                               data_prepare/SegFix_offset_helper.py:97-132 (constant stencil table)
    direction_one_hot                   train_util_dam.py:123-142 (inline block, wrapped)
    LabelEncoding                       my_transforms.py:661-837, the transform WITHOUT direction targets
                                        (cdnet_b200.api.LabelEncoding is my_transforms_direction's)

Same names, argument meaning, return types and error behaviour as the reference; every computation runs in the
sm_100a kernels of csrc/training.cu behind the C ABI (include/cdnet_b200.h).  There is no CPU fallback.
"""
import os

import numpy as np
import torch

from . import _cabi
from ._cabi import CdnetError, check
from .api import _device, _workspace, _stream, _ptr

_ALIGN_CLASSES = (4, 8, 16, 32)
_VECTOR_CLASSES = (4, 5, 8, 9, 16, 17, 32)


class DTOffsetConfig(object):
    """data_prepare/SegFix_offset_helper.py:21-46 (the fields the hot path reads); frozen at import like the
    reference's."""
    max_distance = int(os.environ.get("dt_max_distance", 5))
    min_distance = int(os.environ.get("dt_min_distance", 0))
    direction_classes = 8
    num_classes = int(os.environ.get("dt_num_classes", direction_classes))
    assert num_classes in (4, 8, 16, 32,)
    c4_align_axis = os.environ.get("c4_align_axis") is not None


def _no_c4_axis():
    if DTOffsetConfig.c4_align_axis:
        raise NotImplementedError("c4_align_axis (SegFix_offset_helper.py:46,53-60) is out of scope")


def _sincos_table(num_classes, use_torch):
    """(sin, cos) of every bin centre, evaluated by the same library call the reference makes
    (np.deg2rad / np.sin on float64, or pi/180*x and torch.sin on float32), plus the row for angle 0.0"""
    if num_classes == 4:
        centres = np.arange(4, dtype=np.float64) * 90.0 - 135.0
    else:
        centres = -180.0 + (360 / num_classes) * np.arange(num_classes, dtype=np.float64)
    centres = np.concatenate([centres, [0.0]])
    if use_torch:
        a = torch.from_numpy(centres).float()
        rad = np.pi / 180.0 * a
        tab = torch.stack([torch.sin(rad), torch.cos(rad)], dim=1).double().numpy()
    else:
        if num_classes == 4:
            centres = centres.astype(np.float32)  # align_angle_c4 hands back float32 (:302-306)
        rad = np.deg2rad(centres)
        tab = np.stack([np.sin(rad), np.cos(rad)], axis=1).astype(np.float64)
    return np.ascontiguousarray(tab)


def _as_float_tensor(a, what):
    """numpy / torch floating array -> contiguous device tensor (float32 or float64)"""
    t = torch.from_numpy(np.ascontiguousarray(a)) if isinstance(a, np.ndarray) else a
    if t.dtype not in (torch.float32, torch.float64):
        if t.dtype in (torch.float16, torch.bfloat16):
            t = t.float()
        else:
            t = t.double()  # integer angles promote like numpy's comparisons do
    return t.to(_device(t.device if t.is_cuda else None)).contiguous()


def _result(t, like, return_tensor):
    if return_tensor:
        return t.to(like.device) if isinstance(like, torch.Tensor) else t
    return t.cpu().numpy()


class Sobel(object):
    """data_prepare/SegFix_offset_helper.py:97-132: the 11 x 11 (ksize x ksize) gradient stencil of the direction
    targets as a torch tensor [2,1,k,k], channel 0 = d/dy (row offset / r^2), channel 1 = d/dx, cached per ksize.
    A table of constants (the kernels of csrc/targets.cu hold the same weights in constant memory)."""
    _caches = {}
    ksize = 11

    @staticmethod
    def _generate_sobel_kernel(shape, axis):
        """axis 0: column offset / r^2, axis 1: row offset / r^2 (float64 quotient stored to float32); centre 0"""
        cj, ci = int((shape[0] - 1) / 2.0), int((shape[1] - 1) / 2.0)
        jj, ii = np.mgrid[-cj:shape[0] - cj, -ci:shape[1] - ci]
        r2 = (ii * ii + jj * jj).astype(np.float64)
        odd = shape[0] % 2 == 1 and shape[1] % 2 == 1
        if odd:
            r2[cj, ci] = 1.0
        k = ((ii if axis == 0 else jj) / r2).astype(np.float32)
        if odd:
            k[cj, ci] = 0.0
        return torch.from_numpy(k).unsqueeze(0)

    @classmethod
    def kernel(cls, ksize=None):
        if ksize is None:
            ksize = cls.ksize
        if ksize not in cls._caches:
            sobel_x, sobel_y = (cls._generate_sobel_kernel((ksize, ksize), i) for i in (0, 1))
            cls._caches[ksize] = torch.cat([sobel_y, sobel_x], dim=0).view(2, 1, ksize, ksize)
        return cls._caches[ksize]


class DTOffsetHelper(object):
    """data_prepare/SegFix_offset_helper.py DTOffsetHelper: the static methods on the geometry hot path."""

    @staticmethod
    def label_to_vector(labelmap, num_classes=DTOffsetConfig.num_classes):
        """:246-261.  labelmap torch integer tensor [N,H,W] -> int64 [N,2,H,W] (dh, dw) on labelmap.device."""
        assert isinstance(labelmap, torch.Tensor)
        _no_c4_axis()
        if num_classes not in _VECTOR_CLASSES:
            raise KeyError(num_classes)  # label_to_vector_mapping[num_classes]
        if labelmap.dim() != 3:
            raise RuntimeError("label_to_vector expects a [N,H,W] tensor (the reference's permute(0,3,1,2))")
        L = _cabi.lib()
        dev = _device(labelmap.device if labelmap.is_cuda else None)
        lab = labelmap
        if lab.dtype not in (torch.uint8, torch.int32, torch.int64):
            if lab.dtype.is_floating_point:
                # `labelmap == idx` on floats matches only exact integers: map the rest to "no class"
                lab = torch.where(lab == lab.round(), lab, torch.full_like(lab, -1)).to(torch.int64)
            else:
                lab = lab.to(torch.int64)
        lab = lab.to(dev).contiguous()
        N, H, W = lab.shape
        out = torch.empty((N, 2, H, W), dtype=torch.int64, device=dev)
        if lab.numel():
            check(L.cdnet_label_to_vector(_ptr(lab), lab.element_size(), _ptr(out), N, H * W, int(num_classes),
                                          _stream()), "cdnet_label_to_vector")
        return out.to(labelmap.device)

    @staticmethod
    def _align(angle_map, num_classes, return_tensor, want_snapped=True, want_index=True):
        if return_tensor:
            assert isinstance(angle_map, torch.Tensor)
        else:
            assert isinstance(angle_map, np.ndarray)
        if num_classes not in _ALIGN_CLASSES:
            raise CdnetError("align_angle: num_classes must be one of %r" % (_ALIGN_CLASSES,))
        if num_classes == 4:
            _no_c4_axis()
        L = _cabi.lib()
        a = _as_float_tensor(angle_map, "angle_map")
        dev = a.device
        # dtype of the snapped angle: np.float on the numpy path, .float() on the torch path and for 4 classes
        sdt = torch.float32 if (return_tensor or num_classes == 4) else torch.float64
        snapped = torch.empty(a.shape, dtype=sdt, device=dev) if want_snapped else None
        index = torch.empty(a.shape, dtype=torch.int64, device=dev) if want_index else None
        if a.numel():
            check(L.cdnet_align_angle(_ptr(a), a.element_size(), _ptr(snapped), 4 if sdt == torch.float32 else 8,
                                      _ptr(index), a.numel(), int(num_classes), _stream()), "cdnet_align_angle")
        return snapped, index

    @staticmethod
    def align_angle_c4(angle_map, return_tensor=False):
        """:286-309."""
        s, i = DTOffsetHelper._align(angle_map, 4, return_tensor)
        return _result(s, angle_map, return_tensor), _result(i, angle_map, return_tensor)

    @staticmethod
    def align_angle(angle_map, num_classes=DTOffsetConfig.num_classes, return_tensor=False):
        """:311-341 -> (snapped angle, bin index).  numpy: (float64, int64); torch: (float32, int64)."""
        s, i = DTOffsetHelper._align(angle_map, num_classes, return_tensor)
        return _result(s, angle_map, return_tensor), _result(i, angle_map, return_tensor)

    @staticmethod
    def angle_to_vector(angle_map, num_classes=DTOffsetConfig.num_classes, return_tensor=False):
        """:423-450 -> [...,2] (sin, cos) of the snapped angle; numpy float64, torch float32."""
        if return_tensor:
            assert isinstance(angle_map, torch.Tensor)
        else:
            assert isinstance(angle_map, np.ndarray)
        if num_classes is None:
            raise NotImplementedError("angle_to_vector without snapping (num_classes=None) is out of scope")
        if num_classes not in _ALIGN_CLASSES:
            raise CdnetError("angle_to_vector: num_classes must be one of %r" % (_ALIGN_CLASSES,))
        if num_classes == 4:
            _no_c4_axis()
        L = _cabi.lib()
        a = _as_float_tensor(angle_map, "angle_map")
        odt = torch.float32 if return_tensor else torch.float64
        vec = torch.empty(tuple(a.shape) + (2,), dtype=odt, device=a.device)
        tab = _sincos_table(int(num_classes), bool(return_tensor))
        if a.numel():
            check(L.cdnet_angle_to_vector(_ptr(a), a.element_size(), _ptr(vec), vec.element_size(), tab.ctypes.data,
                                          a.numel(), int(num_classes), _stream()), "cdnet_angle_to_vector")
        return _result(vec, angle_map, return_tensor)

    @staticmethod
    def angle_to_direction_label(angle_map, seg_label_map=None, distance_map=None,
                                 num_classes=DTOffsetConfig.num_classes, extra_ignore_mask=None,
                                 return_tensor=False):
        """:452-484: bin index, `num_classes` beyond max_distance, -1 where ignored."""
        if return_tensor:
            assert isinstance(angle_map, torch.Tensor)
            assert isinstance(seg_label_map, torch.Tensor) or seg_label_map is None
        else:
            assert isinstance(angle_map, np.ndarray)
            assert isinstance(seg_label_map, np.ndarray) or seg_label_map is None
        _, label = DTOffsetHelper._align(angle_map, num_classes, return_tensor, want_snapped=False)

        def dev(x):
            t = torch.from_numpy(np.ascontiguousarray(x)) if isinstance(x, np.ndarray) else x
            return t.to(label.device)
        if distance_map is not None:
            label[dev(distance_map) > DTOffsetConfig.max_distance] = num_classes
        if seg_label_map is not None:
            label[dev(seg_label_map) == -1] = -1
        if extra_ignore_mask is not None:
            label[dev(extra_ignore_mask).bool()] = -1
        return _result(label, angle_map, return_tensor)

    @staticmethod
    def vector_to_label(vector_map, num_classes=DTOffsetConfig.num_classes, return_tensor=False):
        """:486-506: [...,2] (v0, v1) -> int64 bin index of degrees(arctan2(v0, v1))."""
        if return_tensor:
            assert isinstance(vector_map, torch.Tensor)
        else:
            assert isinstance(vector_map, np.ndarray)
        if num_classes not in _ALIGN_CLASSES:
            raise CdnetError("vector_to_label: num_classes must be one of %r" % (_ALIGN_CLASSES,))
        if num_classes == 4:
            _no_c4_axis()
        L = _cabi.lib()
        v = _as_float_tensor(vector_map, "vector_map")
        if v.shape[-1] != 2:
            raise IndexError("vector_map[..., 1]: the last axis must hold (v0, v1)")
        label = torch.empty(v.shape[:-1], dtype=torch.int64, device=v.device)
        if label.numel():
            check(L.cdnet_vector_to_label(_ptr(v), v.element_size(), _ptr(label), label.numel(), int(num_classes),
                                          _stream()), "cdnet_vector_to_label")
        return _result(label, vector_map, return_tensor)


# =====================================================================================================
# direction one-hot + foreground mask (train_util_dam.py:123-142)
# =====================================================================================================
def direction_one_hot_cuda(target_direction, target, direction_classes):
    """target_direction int64 [B,H,W] class ids, target [B,H,W] ternary {0,1,2} (uint8 or int64), both CUDA ->
    (float32 [B,C,H,W], status int32 [B]).  Like the reference (:139) the foreground mask of EVERY tile is
    `target[0]`."""
    L = _cabi.lib()
    dev = _device(target_direction.device)
    d = target_direction.to(torch.int64).contiguous()
    t = target
    if t.dtype not in (torch.uint8, torch.int64):
        t = t.to(torch.int64)
    t = t.contiguous()
    B, H, W = d.shape
    assert t.dim() == 3 and tuple(t.shape[1:]) == (H, W)
    C = int(direction_classes)
    out = torch.empty((B, C, H, W), dtype=torch.float32, device=dev)
    status = torch.empty((B,), dtype=torch.int32, device=dev)
    nb = L.cdnet_direction_one_hot_workspace_bytes(B)
    ws = _workspace(nb, dev)
    check(L.cdnet_direction_one_hot(_ptr(d), _ptr(t[0]), t.element_size(), _ptr(out), _ptr(status), B, C, H * W,
                                    _ptr(ws), ws.numel(), _stream()), "cdnet_direction_one_hot")
    return out, status


def direction_one_hot(target_direction0, target, direction_classes):
    """train_util_dam.py:123-142 as a function: target_direction0 [B,H,W] integer tensor, target [B,H,W] ternary
    tensor -> float32 [B,direction_classes,H,W] on target_direction0's device (the reference builds it on the CPU
    and uploads it afterwards; hand in CUDA tensors to keep everything resident)."""
    assert isinstance(target_direction0, torch.Tensor) and target_direction0.dim() == 3
    dev = _device(target_direction0.device if target_direction0.is_cuda else None)
    out, status = direction_one_hot_cuda(target_direction0.to(dev), target.to(dev), direction_classes)
    if int((status & _cabi.S_CLASS_RANGE).max()):
        raise IndexError("index is out of bounds for dimension 1 with size %d" % int(direction_classes))
    return out.to(target_direction0.device)


# =====================================================================================================
# my_transforms.LabelEncoding (no direction targets)
# =====================================================================================================
def ternary_label_cuda(ch0, mode, ch1=None):
    """ch0 (and ch1 for mode 3) uint8 [B,H,W] CUDA -> uint8 [B,H,W] in {0,127,255} (include/cdnet_b200.h)."""
    L = _cabi.lib()
    dev = _device(ch0.device)
    a = ch0.contiguous()
    b = ch1.contiguous() if ch1 is not None else None
    assert a.dtype == torch.uint8 and (b is None or (b.dtype == torch.uint8 and b.shape == a.shape))
    B, H, W = a.shape
    out = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
    check(L.cdnet_ternary_label(_ptr(a), _ptr(b), int(mode), _ptr(out), B, H, W, _stream()), "cdnet_ternary_label")
    return out


class LabelEncoding(object):
    """Drop-in for my_transforms.LabelEncoding (my_transforms.py:661-837) with do_direction = 0, the form
    options.py builds for the models without a direction branch: `LabelEncoding(out_c, radius, do_direction)(imgs)`
    replaces imgs[2] by the PIL 'L' label image {0,127,255}.

    do_direction = 1 (my_transforms.py:763-836; dead in the reference's own flows, which pair this module with
    direction = 0 models, train.py:77-81) is built for out_c = 3: instances = `measure.label` of the label (not
    dilated, no watershed), nucleus centre = `peak_local_max(distance_transform_edt(nucleus), exclude_border=0,
    num_peaks=1)` = the first raster maximum of the nucleus's own distance transform -- scikit-image's order among
    EQUAL maxima is unpinned (numpy's unstable argsort) -- then the same distance-to-centre / Sobel / quantiser chain
    as the direction-aware transform.  Returns (img, weight, PIL ternary, float16 point map, int64 direction classes).
    The number of direction classes is env `dt_num_classes` (default 8) unless `num_classes` is given."""

    def __init__(self, out_c=3, radius=1, do_direction=0, num_classes=None):
        import os
        self.out_c = out_c
        self.radius = 1  # the reference ignores its argument (:668)
        self.do_direction = do_direction
        self.num_classes = int(os.environ.get("dt_num_classes", 8)) if num_classes is None else int(num_classes)
        if do_direction == 1 and out_c != 3:
            raise NotImplementedError("my_transforms.LabelEncoding with do_direction=1 is built for out_c=3 only")

    @staticmethod
    def _u8(a):
        a = np.asarray(a)
        if a.dtype != np.uint8:
            if a.size and (a.min() < 0 or a.max() > 255):
                raise CdnetError("label values must fit uint8, as data_folder.py:29,37 delivers them")
            a = a.astype(np.uint8)
        return np.ascontiguousarray(a)

    def encode(self, label):
        """label image ([H,W] or [H,W,C]) -> uint8 [H,W] in {0,127,255}"""
        if not isinstance(label, np.ndarray):
            label = np.array(label)
        two_d = label.ndim == 2
        ch0 = self._u8(label if two_d else label[:, :, 0])
        # `len(np.unique(label_inside))`, my_transforms.py:681-686 (np.unique(label) of the whole array for 2-D input)
        from .api import label_stats_cuda
        dev = _device()
        d0 = torch.from_numpy(ch0).to(dev)[None]
        instance_level = int(label_stats_cuda(d0)[0][0]) > 2
        ch1 = None
        if self.out_c != 3:
            if two_d:
                raise IndexError("too many indices for array: array is 2-dimensional, but 3 were indexed")
            if instance_level:
                mode = 2
            else:
                mode = 3
                if label.shape[2] < 2:
                    raise IndexError("index 1 is out of bounds for axis 2 with size %d" % label.shape[2])
                ch1 = torch.from_numpy(self._u8(label[:, :, 1])).to(dev)[None]
        else:
            mode = 0 if instance_level else 1
        return ternary_label_cuda(d0, mode, ch1)[0].cpu().numpy()

    def encode_direction(self, label):
        """label image -> (ternary uint8, point float16, direction int64) of do_direction = 1 (out_c = 3)"""
        if not isinstance(label, np.ndarray):
            label = np.array(label)
        ch0 = self._u8(label if label.ndim == 2 else label[:, :, 0])
        from .api import encode_targets_cuda, label_stats_cuda
        dev = _device()
        d0 = torch.from_numpy(ch0).to(dev)[None]
        instance_level = int(label_stats_cuda(d0)[0][0]) > 2
        tern, point, direction = encode_targets_cuda(d0, instance_level=4 if instance_level else 5,
                                                     num_classes=self.num_classes)
        return tern[0].cpu().numpy(), point[0].cpu().numpy(), direction[0].cpu().numpy()

    def __call__(self, imgs):
        from PIL import Image
        out_imgs = list(imgs)
        if self.do_direction == 1:
            tern, point, direction = self.encode_direction(imgs[2])
            out_imgs[2] = Image.fromarray(tern)
            out_imgs.append(point)
            out_imgs.append(direction)
            return tuple(out_imgs)
        out_imgs[2] = Image.fromarray(self.encode(imgs[2]))
        return tuple(out_imgs)


from .api import guard_public_functions as _guard  # noqa: E402  (device guard, see api._on_tensor_device)
_guard(globals())
