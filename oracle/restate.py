"""oracle/restate.py -- CPU restatement of CDNet's geometry hot path.  TEST INFRASTRUCTURE ONLY.

Every function cites the reference file:line it follows (paths relative to /root/reference).
It restates the reference's algorithm with numpy / scipy.ndimage / numba (the libraries the
reference itself calls) plus two C helpers (oracle/ws_flood.c) for the pieces whose third-party
implementation is absent from this image or not portable across hosts:

  * scikit-image (unpinned, un-vendored): `dilation/erosion/disk`,
    `remove_small_objects`, `measure.label`, `watershed` are restated from their published
    semantics.  PARITY UNPINNED for: watershed tie-breaking between age-0 marker pixels (this
    build's canonical order is (value, age, raster index)), the `selem=None` default footprint
    (taken as the connectivity-1 cross) and `measure.label` numbering (raster-first).
  * torch CPU `conv2d` (f32) is restated as the sequential FMA chain it was measured to equal.

Pinned against: the reference itself executed verbatim in the build container
(`oracle/ref_loader.py`; tests/test_oracle_vs_reference.py) and the golden vectors generated from
it (`oracle/make_goldens.py` -> tests/golden/*.npz; tests/test_oracle_golden.py).  The reference
ships no tests, fixtures or known-answer vectors of its own.

The per-instance loops of the reference are kept (`literal=True`, O(N_inst*H*W), used as the CPU
baseline in bench.py); `literal=False` restricts each instance's work to its bounding window,
which is exact (zero taps are exact no-ops of an FMA chain; the EDT of a 4-connected component
equals the global EDT) and lets 1000x1000 tiles finish in seconds.
"""
import math

import numpy as np
from scipy import ndimage as ndi

from . import clib

# --------------------------------------------------------------------------------------------
# class index -> (dh, dw) vectors, data_prepare/SegFix_offset_helper.py:50-89 (keys 5, 9, 17)
# --------------------------------------------------------------------------------------------
_RING8 = [(0, -1), (-1, -1), (-1, 0), (-1, 1), (0, 1), (1, 1), (1, 0), (1, -1)]
_RING16 = [(0, -2), (-1, -2), (-2, -2), (-2, -1), (-2, 0), (-2, 1), (-2, 2), (-1, 2),
           (0, 2), (1, 2), (2, 2), (2, 1), (2, 0), (2, -1), (2, -2), (1, -2)]
CLASS_VECTORS = {
    5: [(0, 0), (-1, -1), (-1, 1), (1, 1), (1, -1)],
    9: [(0, 0)] + _RING8,
    17: [(0, 0)] + _RING16,
}


def circshift(matrix_ori, direction, shiftnum1, shiftnum2):
    """data_prepare/getDirectionDiffMap.py:14-42 -- zero-filled shift of [C,H,W].
    direction 1/2 move content up by shiftnum1, 3/4 move it down; 1/3 move it left by shiftnum2,
    2/4 move it right."""
    m = np.asarray(matrix_ori)
    c, h, w = m.shape
    out = np.zeros_like(m)
    s1, s2 = int(shiftnum1), int(shiftnum2)
    if direction in (1, 2):
        ys, yd = slice(s1, h), slice(0, h - s1)
    else:
        ys, yd = slice(0, h - s1), slice(s1, h)
    if direction in (1, 3):
        xs, xd = slice(s2, w), slice(0, w - s2)
    else:
        xs, xd = slice(0, w - s2), slice(s2, w)
    out[:, yd, xd] = m[:, ys, xs]
    return out


def label_to_vector(labelmap, num_classes):
    """SegFix_offset_helper.py:246-261 -> int64 [2,H,W] (dh, dw); unknown classes -> (0,0)."""
    lab = np.asarray(labelmap)
    table = np.zeros((256, 2), dtype=np.int64)
    vec = np.asarray(CLASS_VECTORS[num_classes], dtype=np.int64)
    table[:len(vec)] = vec
    idx = np.where((lab >= 0) & (lab < len(vec)), lab, 255).astype(np.int64)
    return np.moveaxis(table[idx], -1, 0)


def generate_dd_map(label_direction, direction_classes):
    """data_prepare/getDirectionDiffMap.py:44-108 -- direction-difference map, f32 [H,W].

    cos between each pixel's class vector and its 8 (9/17 classes) or 4 axial (5 classes)
    zero-padded neighbours, f64 -> f32 (:90-97); min over `direction_classes-1` channels of which
    only 8 (or 4) are filled (:90 vs :69-88, so 17 classes get min(.,0)); background -> 1 (:101);
    1 - around (:103); min-max normalise per image, 0/0 -> NaN (:104-106)."""
    lab = np.asarray(label_direction)
    H, W = lab.shape
    v = label_to_vector(lab, direction_classes)  # [2,H,W] int64
    if direction_classes - 1 == 4:
        shifts = [(1, 1, 0), (3, 0, 1), (4, 0, 1), (3, 1, 0)]
    elif direction_classes - 1 in (8, 16):
        shifts = [(1, 1, 1), (1, 1, 0), (2, 1, 1), (3, 0, 1), (4, 0, 1), (3, 1, 1), (3, 1, 0),
                  (4, 1, 1)]
    else:
        shifts = []
    cosv = np.zeros((H, W, direction_classes - 1), dtype=np.float32)
    norm_self = np.sqrt((v[0] ** 2 + v[1] ** 2).astype(np.float64))
    for k, (d, s1, s2) in enumerate(shifts):
        u = circshift(v, d, s1, s2)
        num = v[0] * u[0] + v[1] * u[1]
        den = norm_self * np.sqrt((u[0] ** 2 + u[1] ** 2).astype(np.float64)) + 0.000001
        cosv[:, :, k] = num / den
    m = cosv.min(axis=2)
    m[lab == 0] = 1
    d = 1 - np.around(m)
    dmax, dmin = d.max(), d.min()
    with np.errstate(invalid="ignore", divide="ignore"):
        return (d - dmin) / (dmax - dmin)


def ddm_codes(label_direction, direction_classes):
    """the un-normalised d in {0,1,2} of generate_dd_map (uint8) -- Appendix B.1 LUT form."""
    lab = np.asarray(label_direction)
    vec = np.asarray(CLASS_VECTORS[direction_classes], dtype=np.int64)
    n = len(vec)
    dot = vec @ vec.T
    nrm = np.sqrt((vec ** 2).sum(1).astype(np.float64))
    lut = np.around((dot / (nrm[:, None] * nrm[None, :] + 0.000001)).astype(np.float32))
    full = np.zeros((256, 256), dtype=np.float32)
    full[:n, :n] = lut
    p = np.pad(np.where(lab < n, lab, 255).astype(np.int64), 1)
    H, W = lab.shape
    offs = ([(1, 0), (0, 1), (0, -1), (-1, 0)] if direction_classes == 5 else
            [(dy, dx) for dy in (-1, 0, 1) for dx in (-1, 0, 1) if (dy, dx) != (0, 0)])
    m = np.full((H, W), 1.0, dtype=np.float32)
    c = p[1:-1, 1:-1]
    for dy, dx in offs:
        m = np.minimum(m, full[c, p[1 + dy:H + 1 + dy, 1 + dx:W + 1 + dx]])
    if direction_classes == 17:
        m = np.minimum(m, 0)
    m[lab == 0] = 1
    return (1 - m).astype(np.uint8)


# --------------------------------------------------------------------------------------------
# quantiser, SegFix_offset_helper.py:311-341, 423-450, 486-506
# --------------------------------------------------------------------------------------------
def align_angle(angle_map, num_classes=8):
    """SegFix_offset_helper.py:311-341 -> (snapped angle f64, class index int64); upper-inclusive
    bins centred on -180 + k*step; num_classes 4 follows align_angle_c4 (:286-309)."""
    a = np.asarray(angle_map)
    if num_classes == 4:
        idx = np.clip(np.trunc((a + 180) / 90).astype(np.int64), 0, 3)
        return (idx * 90 - 135).astype(np.float32), idx
    step = 360 / num_classes
    snapped = np.zeros(a.shape, dtype=np.float64)
    idx = np.zeros(a.shape, dtype=np.int64)
    wrap = (a <= (-180 + step / 2)) | (a > (180 - step / 2))
    snapped[wrap] = -180
    for i in range(1, num_classes):
        mid = -180 + step * i
        sel = (a > (mid - step / 2)) & (a <= (mid + step / 2))
        snapped[sel] = mid
        idx[sel] = i
    return snapped, idx


def angle_to_vector(angle_map, num_classes=8):
    """SegFix_offset_helper.py:423-450 -> f64 [...,2] = (sin, cos) of the snapped angle."""
    snapped, _ = align_angle(np.asarray(angle_map), num_classes)
    r = np.deg2rad(snapped)
    return np.stack([np.sin(r), np.cos(r)], axis=-1).astype(np.float64)


def vector_to_label(vector_map, num_classes=8):
    """SegFix_offset_helper.py:486-506 (+452-484 with no ignore mask) -> int64 class index."""
    ang = np.rad2deg(np.arctan2(vector_map[..., 0], vector_map[..., 1]))
    return align_angle(ang, num_classes)[1]


def sobel_kernels(ksize=11):
    """SegFix_offset_helper.py:102-132 -> f32 [2,k,k]; [0] = d/dy (row offset / r^2), [1] = d/dx."""
    c = (ksize - 1) // 2
    jj, ii = np.mgrid[-c:c + 1, -c:c + 1]
    r2 = (ii * ii + jj * jj).astype(np.float64)
    r2[c, c] = 1.0
    ky = (jj / r2).astype(np.float32)
    kx = (ii / r2).astype(np.float32)
    ky[c, c] = 0
    kx[c, c] = 0
    return np.stack([ky, kx])


# --------------------------------------------------------------------------------------------
# scikit-image / scipy morphology semantics
# --------------------------------------------------------------------------------------------
CROSS = ndi.generate_binary_structure(2, 1)
FULL3 = np.ones((3, 3), dtype=bool)


def disk(r):
    y, x = np.mgrid[-r:r + 1, -r:r + 1]
    return (x * x + y * y) <= r * r


def dilate(img, footprint=CROSS):
    """skimage.morphology.dilation == ndi.grey_dilation(footprint), mode reflect."""
    return ndi.grey_dilation(np.asarray(img), footprint=np.asarray(footprint)[::-1, ::-1])


def erode(img, footprint=CROSS):
    return ndi.grey_erosion(np.asarray(img), footprint=np.asarray(footprint))


def remove_small_objects(ar, min_size):
    """skimage.morphology.remove_small_objects: bool input -> 4-connected components; integer
    input -> the values are the labels; drop those with count < min_size; no renumbering."""
    ar = np.asarray(ar)
    out = ar.copy()
    if min_size == 0:
        return out
    ccs = ndi.label(ar, CROSS)[0] if ar.dtype == bool else out
    sizes = np.bincount(ccs.ravel())
    out[(sizes < min_size)[ccs]] = 0
    return out


def label8(x):
    """skimage.measure.label on a binary image: 8-connected, raster-first ids, int64."""
    return ndi.label(np.asarray(x) != 0, FULL3)[0].astype(np.int64)


def label4(x):
    """scipy.ndimage.label default structure: 4-connected, raster-first ids, int32."""
    return ndi.label(np.asarray(x) != 0)[0]


# --------------------------------------------------------------------------------------------
# postproc_other.process, postproc_other.py:15-54 (ws branch and the unet/micronet no-ws head)
# --------------------------------------------------------------------------------------------
def inst_dist_map(labels, literal=True):
    """postproc_other.py:16-27 gen_inst_dst_map: per 4-connected instance EDT scaled to
    uint8(255*d/max d), truncating."""
    labels = np.asarray(labels)
    canvas = np.zeros(labels.shape, dtype=np.uint8)
    if literal:
        for k in np.unique(labels):
            if k == 0:
                continue
            d = ndi.distance_transform_edt(labels == k)
            canvas += (255 * (d / np.amax(d))).astype(np.uint8)
        return canvas
    if labels.max() == 0:
        return canvas
    if (labels > 0).all():
        return inst_dist_map(labels, literal=True)
    d = ndi.distance_transform_edt(labels > 0)  # Appendix B.2: == per-instance EDT for 4-conn comps
    mx = ndi.maximum(d, labels, index=np.arange(labels.max() + 1))
    mx[0] = 1.0
    return np.where(labels > 0, (255 * (d / mx[labels])), 0).astype(np.uint8)


def process(pred, model_mode="modelName", min_size=10, ws=True, literal=True, order="stable",
            return_parts=False):
    """postproc_other.py:15-54.  Mutates `pred` in place (binarise at 0.5, :31-32) like the
    reference.  Returns int32 labels with id gaps."""
    if model_mode == "dcan":
        raise NotImplementedError("dcan branch (postproc_other.py:69-97) is out of scope")
    assert len(pred.shape) == 2, "Prediction shape is not HW"
    hi = pred > 0.5
    pred[hi] = 1
    pred[~hi] = 0
    if model_mode in ("unet", "micronet"):
        ws = False
    if not ws:
        out = remove_small_objects(label4(ndi.binary_fill_holes(pred)), min_size)
        if model_mode == "micronet":
            raise NotImplementedError("micronet tail (postproc_other.py:56-68) is out of scope")
        return out
    comps = label4(pred)
    if comps.size == 0 or comps.min() != 0:
        # postproc_other.py:18-19: `nuc_list = list(np.unique(ann)); nuc_list.remove(0)` -- a mask without
        # any background pixel makes list.remove raise
        raise ValueError("list.remove(x): x not in list")
    dist = inst_dist_map(comps, literal=literal)
    marker = dist > 125
    marker = ndi.binary_fill_holes(marker)
    marker = ndi.binary_erosion(marker, iterations=1)
    marker = remove_small_objects(label4(marker), min_size)
    neg = (-dist)  # uint8 modular negation, postproc_other.py:47
    flooded = clib.ws_flood(neg, marker * (pred != 0), pred != 0, order=order)
    out = remove_small_objects(flooded, min_size)
    if return_parts:
        return out, {"comps": comps, "dist": dist, "marker": marker, "flooded": flooded}
    return out


# --------------------------------------------------------------------------------------------
# inference post-processing blocks
# --------------------------------------------------------------------------------------------
def dam_postprocess(prob_maps, point_maps, dcm_tta, direction_classes=9, min_area=20, radius=2,
                    postproc=0, model_name="modelName", literal=True, order="stable", voting_first=False):
    """test_dam.py:455-563 with its hard-wired switches (dcm_combined=1, voting_firt=0,
    DDM_switch=100, mseloss=1, direction=1).  prob_maps f32 [3,H,W] (channel 2 is overwritten in
    place, :536), point_maps f32 [1,H,W], dcm_tta 8 maps [8,H,W].
    Returns dict(pred_labeled, pred_inside, pred2, ddm_mean)."""
    prob_maps = np.asarray(prob_maps)
    H, W = prob_maps.shape[1:]
    if voting_first:
        # test_dam.py:471-477 with `voting_firt = 1`: vote the 8 TTA maps, then ONE float32 DDM, no mean
        voted = dcm_voting2(np.stack([np.asarray(m) for m in dcm_tta], axis=2).astype(np.uint8))
        ddm = generate_dd_map(voted, direction_classes)
    elif len(dcm_tta) == 1:
        # single-map variant, test_dam.py:499-502 (the `dcm_combined != 1` branch): float32 DDM, no mean
        ddm = generate_dd_map(np.asarray(dcm_tta[0]).astype(np.uint8), direction_classes)
    else:
        stack = np.zeros((H, W, 8), dtype=np.float64)
        for t in range(8):
            stack[:, :, t] = generate_dd_map(np.asarray(dcm_tta[t]).astype(np.uint8), direction_classes)
        ddm = np.mean(stack, axis=2)  # :489
    with np.errstate(invalid="ignore", divide="ignore"):
        gate = (point_maps[0] / np.max(point_maps) > 0.2) * 1  # :530
    gate = dilate(gate, disk(1))  # :531
    eb = 2 * (ddm - ddm * gate)  # :532-534
    assert np.min(eb) >= 0  # :535 (fails on a NaN DDM, i.e. a constant direction map)
    prob_maps[2, :, :] = (prob_maps[2, :, :] + 0.5 * eb) * (1 + eb)  # :536
    pred = np.argmax(prob_maps, axis=0)
    inside = pred == 1
    filled = ndi.binary_fill_holes(inside)  # :546
    pred2 = remove_small_objects(filled, min_area).astype(np.uint8)  # :548,554
    if int(postproc) == 1:
        lab = process(inside.astype(np.uint8) * 255, model_name, literal=literal, order=order)
    else:
        lab = label8(pred2)  # :561
    lab = dilate(lab, disk(radius))  # :563
    return {"pred_labeled": lab, "pred_inside": inside, "pred2": pred2, "ddm_mean": ddm}


def plain_postprocess(prob_maps, min_area=20, radius=2, postproc=0, model_name="modelName",
                      multi_class=True, literal=True, order="stable"):
    """test.py:270-295."""
    prob_maps = np.asarray(prob_maps)
    inside = (np.argmax(prob_maps, axis=0) == 1) if multi_class else (prob_maps[0] >= 0.5)
    filled = ndi.binary_fill_holes(inside)
    pred2 = remove_small_objects(filled, min_area).astype(np.uint8)
    if int(postproc) == 1:
        lab = process(inside.astype(np.uint8) * 255, model_name, min_size=min_area,
                      literal=literal, order=order)
    else:
        lab = label8(pred2)
    lab = dilate(lab, disk(radius))
    return {"pred_labeled": lab, "pred_inside": inside, "pred2": pred2}


def dcm_voting2(direct_map):
    """utils.py:1150-1159 -- remap the 8 TTA direction maps into the un-flipped frame and vote
    (first maximum wins).  direct_map u8 [H,W,8] -> int64 [H,W]."""
    ring = np.arange(1, 9)
    perms = [ring, np.roll(ring[::-1], 5), np.roll(ring[::-1], 1), np.roll(ring, -4),
             np.roll(ring, -2), np.roll(ring[::-1], 7), np.roll(ring[::-1], 3), np.roll(ring, -6)]
    dm = np.asarray(direct_map)
    H, W = dm.shape[:2]
    votes = np.zeros((H, W, 9), dtype=np.uint8)
    yy, xx = np.mgrid[0:H, 0:W]
    for t in range(8):
        table = np.concatenate([[0], perms[t]])
        cls = dm[:, :, t]
        ok = cls <= 8
        np.add.at(votes, (yy[ok], xx[ok], table[cls[ok]]), 1)
    return np.argmax(votes, axis=2)


# --------------------------------------------------------------------------------------------
# target transform, my_transforms_direction.py:651-885
# --------------------------------------------------------------------------------------------
_RAYS = np.array([(math.sin(2 * math.pi / 8 * k), math.cos(2 * math.pi / 8 * k)) for k in range(8)],
                 dtype=np.float64)

_centre_scan = None


def _get_centre_scan():
    global _centre_scan
    if _centre_scan is None:
        import numba

        @numba.njit(cache=False)
        def scan(win, oy, ox, n, m, rays):
            # win: window of the one-nucleus mask whose top-left pixel is (oy, ox) of the n x m canvas
            best = -1.0
            by = -1
            bx = -1
            h, w = win.shape
            for wy in range(h):
                for wx in range(w):
                    if win[wy, wx] > 0:
                        i = wy + oy  # absolute coordinates: half-to-even rounding depends on them
                        j = wx + ox
                        far = 0.0
                        near = 10000000.0
                        for k in range(8):
                            lo = 0.0
                            hi = 1000.0
                            for _ in range(30):
                                mid = (lo + hi) / 2
                                py = round(i + rays[k, 0] * mid)
                                px = round(j + rays[k, 1] * mid)
                                hit = False
                                if py >= 0 and py < n and px >= 0 and px < m:
                                    qy = py - oy
                                    qx = px - ox
                                    if qy >= 0 and qy < h and qx >= 0 and qx < w and win[qy, qx] > 0:
                                        hit = True
                                if hit:
                                    lo = mid
                                else:
                                    hi = mid
                            far = max(far, hi)
                            near = min(near, hi)
                        c = near / far
                        if c > best:
                            best = c
                            by = i
                            bx = j
            return by, bx

        _centre_scan = scan
    return _centre_scan


def get_centerpoint2(mask, n=None, m=None, window=None, windowed=False):
    """my_transforms_direction.py:651-685 -- pixel of maximum centerness (min/max over 8 rays of the
    30-step bisected reach), first raster-order maximum.  `window=(y0,y1,x0,x1)` restricts the scan
    to a window of the n x m canvas known to contain the whole nucleus (exact); with
    `windowed=True` `mask` already IS that window."""
    mask = np.asarray(mask)
    n = mask.shape[0] if n is None else n
    m = mask.shape[1] if m is None else m
    y0, y1, x0, x1 = (0, n, 0, m) if window is None else window
    win = mask if windowed else mask[y0:y1, x0:x1]
    win = np.ascontiguousarray(win).astype(np.int64)
    by, bx = _get_centre_scan()(win, y0, x0, n, m, _RAYS)
    return [int(by), int(bx)]


def ternary_and_instances(label, literal=True, order="stable", out_c=3):
    """my_transforms_direction.py:714-782 for out_c == 3.
    Returns (new_label u8 {0,1,2}, inside u8 {0,1}, label_instance int, instance_level bool)."""
    label = np.asarray(label)
    inside_src = label if label.ndim == 2 else label[:, :, 0]
    instance_level = len(np.unique(inside_src)) > 2
    new_label = np.zeros(inside_src.shape, dtype=np.uint8)
    if out_c != 3:
        # my_transforms_direction.py:721-739: no boundary class, instances are NOT dilated, inside holds 2s
        if instance_level:
            # measure.label(measure.label(label)[:, :, 0]) == the 2-D equal-value 8-connected components of
            # channel 0, whatever the other channels hold (3-D links can only join what the 2-D relabelling
            # separates again)
            inst = label8_values(label[:, :, 0])
            new_label[inst > 0] = 2
        else:
            new_label[label[:, :, 0] > 255 * 0.5] = 2
            new_label[label[:, :, 1] > 255 * 0.5] = 2
            new_label = erode(new_label, disk(1))  # :733
            inst = label8(new_label)
        return new_label, new_label.copy(), inst, instance_level
    if instance_level:
        new_label[inside_src > 0] = 1
        new_label = remove_small_objects(new_label, 5)  # :746 (value 1 treated as ONE label)
        inside = new_label.copy()
        boun = dilate(inside_src) & (~erode(inside_src, disk(1)))  # :750
        new_label[boun > 0] = 2
        interior = (new_label == 1).astype(np.uint8)
        inst = process(interior * 255, "modelName", min_size=5, literal=literal, order=order)  # :759
        inst = dilate(inst, disk(1))  # :760
    else:
        new_label[inside_src > 255 * 0.5] = 1  # :765
        inside = new_label.copy()
        boun = dilate(new_label) & (~erode(new_label, disk(1)))  # :768
        new_label[boun > 0] = 2
        interior = (new_label == 1).astype(np.uint8)
        inst = dilate(label8(interior), disk(1))  # :773-774
    return new_label, inside, inst, instance_level


def label_encoding(label, out_c=3, radius=1, do_direction=1, num_classes=8, literal=True,
                   order="stable", conv="fma", return_parts=False):
    """LabelEncoding.__call__, my_transforms_direction.py:697-885.

    Returns (ternary u8 {0,127,255}, point f16 [H,W], direction int64 [H,W]); the last two are None
    when do_direction != 1.  `num_classes` plays the role of env `dt_num_classes`
    (SegFix_offset_helper.py:37-39)."""
    new_label, inside, inst, _ = ternary_and_instances(label, literal=literal, order=order, out_c=out_c)
    ternary = (new_label / 2 * 255).astype(np.uint8)  # :781
    if do_direction != 1:
        return ternary, None, None
    H, W = new_label.shape
    dir_map = np.zeros((H, W, 2), dtype=np.float32)
    label_point = np.zeros((H, W), dtype=np.float64)
    dc_sum = np.zeros((H, W), dtype=np.float64)
    ker = sobel_kernels(11)
    ids = np.unique(inst)[1:]  # :797-800 (assumes a background pixel exists)
    centres = []
    boxes = ndi.find_objects(inst.astype(np.int64)) if not literal else None
    for k in ids:
        if literal:
            y0, y1, x0, x1 = 0, H, 0, W
        else:
            sl = boxes[int(k) - 1]
            y0, y1 = max(0, sl[0].start - 7), min(H, sl[0].stop + 7)
            x0, x1 = max(0, sl[1].start - 7), min(W, sl[1].stop + 7)
        nucleus = (inst[y0:y1, x0:x1] == k).astype(np.int64)
        cy, cx = get_centerpoint2(nucleus, H, W, (y0, y1, x0, x1), windowed=True)  # :813
        centres.append((cy, cx))
        label_point[cy, cx] = 255.0  # :816
        nucleus = dilate(nucleus, disk(1))  # :819
        yy, xx = np.mgrid[y0:y1, x0:x1]
        dist = np.sqrt(((yy - cy) ** 2 + (xx - cx) ** 2).astype(np.float64))  # :820-822 EDT(1-point)
        int_pos = dist * nucleus
        dc = (1 - int_pos / (int_pos.max() + 0.0000001)) * nucleus  # :824
        dc_sum[y0:y1, x0:x1] += dc
        dc32 = dc.astype(np.float32)
        if conv == "fma":
            g = clib.conv11_fma(dc32, ker)  # restates F.conv2d (:827-830)
        else:
            import torch
            g = torch.nn.functional.conv2d(torch.from_numpy(dc32).view(1, 1, *dc32.shape),
                                           torch.from_numpy(ker).view(2, 1, 11, 11),
                                           padding=5)[0].numpy()
        g = np.moveaxis(g, 0, -1).copy()
        g[nucleus == 0, :] = 0  # :832
        sub = dir_map[y0:y1, x0:x1]
        sub[nucleus != 0, :] = 0  # :833
        sub += g  # :834
    assert int(label_point.sum() / 255) == len(ids)  # :836
    point = ndi.gaussian_filter(label_point, sigma=2, order=0).astype(np.float16)  # :842
    angle = np.degrees(np.arctan2(dir_map[:, :, 0], dir_map[:, :, 1]))  # :848 (f32)
    angle[inside == 0] = 0  # :852
    vec = angle_to_vector(angle, num_classes)  # :853
    direction = vector_to_label(vec, num_classes)  # :855
    direction[inside == 0] = -1  # :859-864
    direction = direction + 1
    if return_parts:
        return ternary, point, direction, {"inst": inst, "inside": inside, "dir_map": dir_map,
                                           "centres": np.asarray(centres, dtype=np.int64),
                                           "dc": dc_sum, "angle": angle, "new_label": new_label}
    return ternary, point, direction


# --------------------------------------------------------------------------------------------
# training-side consumers
# --------------------------------------------------------------------------------------------
def direction_one_hot(target_direction0, target, direction_classes):
    """train_util_dam.py:123-142.  target_direction0 int [B,H,W], target ternary [B,H,W] ->
    float32 [B,C,H,W].  The foreground mask of every tile is target[0] (:139, sic); a tile with a
    single distinct direction value gets channel 0 = 1 everywhere (:141)."""
    d = np.asarray(target_direction0)
    t = np.asarray(target)
    B, H, W = d.shape
    out = np.zeros((B, int(direction_classes), H, W), dtype=np.float32)
    fg0 = (t[0] == 1) | (t[0] == 2)
    for j in range(B):
        uniq = np.unique(d[j])
        if len(uniq) > 1:
            for k in uniq:
                if k < 0 or k >= direction_classes:
                    raise IndexError("index %d is out of bounds for dimension 1 with size %d" % (k, direction_classes))
                out[j, k][d[j] == k] = 1
                out[j, k][~fg0] = 0
        else:
            out[j, 0] = 1
    return out


def label_encoding_plain(label, out_c=3):
    """my_transforms.LabelEncoding with do_direction=0 (my_transforms.py:661-761): the ternary /
    binary label image uint8 {0,127,255}."""
    label = np.asarray(label)
    inside = label if label.ndim == 2 else label[:, :, 0]
    instance_level = len(np.unique(inside)) > 2
    new_label = np.zeros(label.shape[:2], dtype=np.uint8)
    if out_c != 3:
        if instance_level:
            new_label[label[:, :, 0] > 0] = 2  # measure.label(label)[:, :, 0] > 0  (:690-696)
        else:
            new_label[label[:, :, 0] > 127.5] = 2
            new_label[label[:, :, 1] > 127.5] = 2
            new_label = erode(new_label, disk(1))  # :703
    elif instance_level:
        inst = label8_values(inside)
        new_label[inst > 0] = 1
        boun = dilate(inst) & (~erode(inst, disk(1)))  # :725
        new_label[boun > 0] = 2
    else:
        new_label[inside > 127.5] = 1
        boun = dilate(new_label) & (~erode(new_label, disk(1)))  # :735
        new_label[boun > 0] = 2
    return (new_label / 2 * 255).astype(np.uint8)


def label_encoding_plain_direction(label, out_c=3, num_classes=8, literal=True, conv="fma", return_parts=False):
    """my_transforms.LabelEncoding with do_direction=1 (my_transforms.py:661-836): like the direction-aware transform
    but on `measure.label(label_inside)` (8-connected components of equal value, NOT dilated, no watershed), and with
    the nucleus centre taken as `peak_local_max(distance_transform_edt(nucleus), exclude_border=0, num_peaks=1)`
    (:775) = the first raster maximum of the nucleus's own distance transform (tie order UNPINNED, see
    oracle/refshim/skimage/feature.py).  Returns (ternary u8, point f16, direction int64)."""
    assert out_c == 3, "restated for the three-class output only"
    label = np.asarray(label)
    label_inside = label if label.ndim == 2 else label[:, :, 0]
    instance_level = len(np.unique(label_inside)) > 2
    ternary = label_encoding_plain(label, out_c)
    inst = label8_values(label_inside)  # :713 / :741 (measure.label of the ids resp. of the {0,255} label)
    inside = (inst > 0) if instance_level else (label_inside > 127.5)  # new_label_inside (:717 / :733)
    H, W = inst.shape
    dir_map = np.zeros((H, W, 2), dtype=np.float32)
    label_point = np.zeros((H, W), dtype=np.float64)
    ker = sobel_kernels(11)
    ids = np.unique(inst)[1:]  # :765-768 (takes the first value for background)
    centres = []
    boxes = ndi.find_objects(inst.astype(np.int64)) if not literal else None
    for k in ids:
        if literal:
            y0, y1, x0, x1 = 0, H, 0, W
        else:
            sl = boxes[int(k) - 1]
            y0, y1 = max(0, sl[0].start - 7), min(H, sl[0].stop + 7)
            x0, x1 = max(0, sl[1].start - 7), min(W, sl[1].stop + 7)
        nucleus = (inst[y0:y1, x0:x1] == k).astype(np.int64)
        dist_i = ndi.distance_transform_edt(nucleus)  # :771
        flat = int(np.argmax(dist_i))  # first raster maximum = peak_local_max(..., num_peaks=1) under a stable order
        cy, cx = y0 + flat // (x1 - x0), x0 + flat % (x1 - x0)
        centres.append((cy, cx))
        label_point[cy, cx] = 255.0  # :778
        nucleus = dilate(nucleus, disk(1))  # :781
        yy, xx = np.mgrid[y0:y1, x0:x1]
        dist = np.sqrt(((yy - cy) ** 2 + (xx - cx) ** 2).astype(np.float64))  # :782-784 EDT(1 - point)
        int_pos = dist * nucleus
        dc = (1 - int_pos / (int_pos.max() + 0.0000001)) * nucleus  # :786
        dc32 = dc.astype(np.float32)
        if conv == "fma":
            g = clib.conv11_fma(dc32, ker)  # restates F.conv2d (:792-795)
        else:
            import torch
            g = torch.nn.functional.conv2d(torch.from_numpy(dc32).view(1, 1, *dc32.shape),
                                           torch.from_numpy(ker).view(2, 1, 11, 11), padding=5)[0].numpy()
        g = np.moveaxis(g, 0, -1).copy()
        g[nucleus == 0, :] = 0  # :796
        sub = dir_map[y0:y1, x0:x1]
        sub[nucleus != 0, :] = 0  # :797
        sub += g  # :798
    point = ndi.gaussian_filter(label_point, sigma=2, order=0).astype(np.float16)  # :802
    angle = np.degrees(np.arctan2(dir_map[:, :, 0], dir_map[:, :, 1]))  # :810
    angle[inside == 0] = 0  # :811
    vec = angle_to_vector(angle, num_classes)  # :812
    direction = vector_to_label(vec, num_classes)  # :814
    direction[inside == 0] = -1  # :817-822
    direction = direction + 1
    if return_parts:
        return ternary, point, direction, {"inst": inst, "inside": inside, "dir_map": dir_map,
                                           "centres": np.asarray(centres, dtype=np.int64), "angle": angle}
    return ternary, point, direction


def label8_values(x):
    """skimage.measure.label(x) for a multi-valued image: 8-connected components of EQUAL value,
    background 0, ids in raster order of each component's first pixel."""
    x = np.asarray(x)
    prov = np.zeros(x.shape, dtype=np.int64)
    nxt = 0
    for v in np.unique(x):
        if v == 0:
            continue
        lab, n = ndi.label(x == v, structure=FULL3)
        prov[lab > 0] = lab[lab > 0] + nxt
        nxt += n
    ids, first = np.unique(prov.ravel(), return_index=True)  # first raster pixel of every provisional id
    rank = np.zeros(nxt + 1, dtype=np.int64)
    fg = ids > 0
    rank[ids[fg][np.argsort(first[fg])]] = np.arange(1, int(fg.sum()) + 1)
    return rank[prov]


# --------------------------------------------------------------------------------------------
# test-time augmentation hand-off
# --------------------------------------------------------------------------------------------
def tta_variant_to_original(a, v):
    """test_dam.py:357-367, 426-441 (averaging :445-450): bring the [C,h,w] output of TTA variant v (0 id, 1 hf, 2 vf,
    3 hvf, 4 r90, 5 r90_hf, 6 r90_vf, 7 r90_hvf) back into the frame of the original image."""
    a = np.asarray(a)
    if v & 1:
        a = np.flip(a, 2)
    if v & 2:
        a = np.flip(a, 1)
    if v & 4:
        a = np.rot90(a, k=3, axes=(1, 2))
    return a


def tta_source_index(v, y, x, H, W):
    """closed form of the above: (row, col) in variant v's own frame of original-frame pixel (y, x)"""
    if v < 4:
        return (H - 1 - y if v & 2 else y), (W - 1 - x if v & 1 else x)
    return (x if v & 2 else W - 1 - x), (H - 1 - y if v & 1 else y)


def variant_probmaps(mask_logits, point, dir_logits):
    """get_probmaps, test_dam.py:983-1013 with direction_label=True: float32 softmax over channels
    (max-shifted, channels summed in order), direction[0] *= mask[0], first-maximum argmax."""
    def softmax(z):
        z = np.asarray(z, dtype=np.float32)
        e = np.exp(z - z.max(axis=0, keepdims=True), dtype=np.float32)
        s = np.zeros(e.shape[1:], dtype=np.float32)
        for c in range(e.shape[0]):
            s = s + e[c]
        return e / s
    prob = softmax(mask_logits)
    d = softmax(dir_logits)
    d[0] = d[0] * prob[0]
    return prob, np.asarray(point, dtype=np.float32), np.argmax(d, axis=0)[None]


def tta_merge(mask_logits, point, dir_logits):
    """test_dam.py:299-450: 8 variants -> (prob f32 [3,H,W], point f32 [1,H,W], dcm int64 [8,H,W]);
    probabilities and point maps are summed in variant order and divided by 8 in float32."""
    probs, points, dcms = [], [], []
    for v in range(8):
        p, q, c = variant_probmaps(mask_logits[v], point[v], dir_logits[v])
        probs.append(tta_variant_to_original(p, v))
        points.append(tta_variant_to_original(q, v))
        dcms.append(tta_variant_to_original(c, v)[0])
    prob, pt = probs[0], points[0]
    for v in range(1, 8):
        prob = prob + probs[v]
        pt = pt + points[v]
    return prob / 8, pt / 8, np.stack(dcms)
