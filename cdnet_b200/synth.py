"""Seeded synthetic inputs shaped like the reference's data.

Everything is built from `numpy.random.default_rng(seed)` integer draws and IEEE add/mul/div only
(no libm transcendentals on the data path: trigonometric constants are quantised to 2^-20), so the
same seed produces bit-identical arrays in the build container (where the goldens are generated)
and on the GPU box (where they are checked).  `digest()` lets a test assert exactly that.

* `instance_map`      -- MoNuSeg/CPM17-shaped instance label map (random rotated ellipses, many
                         touching), int32 ids 1..N; `as_uint8_label` wraps ids the way
                         data_folder.py:29,37 delivers them (uint8, 3 channels).
* `postproc_inputs`   -- what test_dam.py feeds its post-processing block (:455-563): 8 TTA
                         direction-argmax maps (uint8), a 3-class probability map (f32 [3,H,W]) and
                         a point map (f32 [1,H,W]).
"""
import hashlib

import numpy as np

_Q = float(1 << 20)


def _q(x):
    return np.round(np.asarray(x, dtype=np.float64) * _Q) / _Q


# 64 quantised rotation angles (cos, sin)
_ROT = np.stack([_q(np.cos(np.arange(64) * (np.pi / 64.0))),
                 _q(np.sin(np.arange(64) * (np.pi / 64.0)))], axis=1)


def digest(*arrays):
    h = hashlib.sha1()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode() + str(a.shape).encode())
        h.update(a.tobytes())
    return h.hexdigest()


def instance_map(seed, H, W, n_target, axes=(6, 16), return_centres=False):
    """int32 [H,W] ids 1..N (0 = background).  Ellipses are laid on a jittered grid in random order
    and only claim free pixels, so neighbours abut (shared borders, diagonal contacts) like
    annotated nuclei do."""
    rng = np.random.default_rng(seed)
    ids = np.zeros((H, W), dtype=np.int32)
    if n_target <= 0:
        return (ids, np.zeros((0, 2), np.int64)) if return_centres else ids
    cell = max(2.0, float(np.sqrt(H * W / float(n_target))))
    gy, gx = int(np.ceil(H / cell)), int(np.ceil(W / cell))
    cells = rng.permutation(gy * gx)
    jit = rng.integers(0, 1 << 16, size=(gy * gx, 2)).astype(np.float64) / float(1 << 16)
    ax = rng.integers(0, 1 << 16, size=(gy * gx, 2)).astype(np.float64) / float(1 << 16)
    rot = rng.integers(0, 64, size=gy * gx)
    lo, hi = float(axes[0]), float(axes[1])
    centres = []
    k = 0
    for c in cells:
        cy = (c // gx + jit[c, 0]) * cell
        cx = (c % gx + jit[c, 1]) * cell
        if cy >= H or cx >= W:
            continue
        a = lo + (hi - lo) * ax[c, 0]
        b = lo + (hi - lo) * ax[c, 1]
        co, si = _ROT[rot[c]]
        r = int(np.ceil(max(a, b))) + 1
        y0, y1 = max(0, int(cy) - r), min(H, int(cy) + r + 1)
        x0, x1 = max(0, int(cx) - r), min(W, int(cx) + r + 1)
        yy = np.arange(y0, y1, dtype=np.float64)[:, None] - cy
        xx = np.arange(x0, x1, dtype=np.float64)[None, :] - cx
        u = (xx * co + yy * si) / a
        v = (yy * co - xx * si) / b
        inside = (u * u + v * v) <= 1.0
        sub = ids[y0:y1, x0:x1]
        claim = inside & (sub == 0)
        if claim.sum() < 12:
            continue
        k += 1
        sub[claim] = k
        centres.append((int(cy), int(cx)))
    if return_centres:
        return ids, np.asarray(centres, dtype=np.int64).reshape(-1, 2)
    return ids


def as_uint8_label(ids):
    """uint8 [H,W,3] as data_folder.py:29,37 delivers an instance label: ids wrapped into 1..255
    (never 0 for a nucleus), replicated over 3 channels."""
    w = np.where(ids > 0, (ids - 1) % 255 + 1, 0).astype(np.uint8)
    return np.repeat(w[:, :, None], 3, axis=2)


def _cross_extreme(a, fn):
    out = a.copy()
    out[1:, :] = fn(out[1:, :], a[:-1, :])
    out[:-1, :] = fn(out[:-1, :], a[1:, :])
    out[:, 1:] = fn(out[:, 1:], a[:, :-1])
    out[:, :-1] = fn(out[:, :-1], a[:, 1:])
    return out


def _shift_zero(a, dy, dx):
    out = np.zeros_like(a)
    H, W = a.shape
    ys, yd = (slice(dy, H), slice(0, H - dy)) if dy >= 0 else (slice(0, H + dy), slice(-dy, H))
    xs, xd = (slice(dx, W), slice(0, W - dx)) if dx >= 0 else (slice(0, W + dx), slice(-dx, W))
    out[yd, xd] = a[ys, xs]
    return out


def _irwin_hall(rng, shape, n=4):
    """approximately normal noise in [0, n), exact in float32"""
    s = rng.integers(0, 256, size=(n,) + tuple(shape), dtype=np.int32).sum(axis=0)
    return (s.astype(np.float32) / np.float32(256.0))


def centripetal_classes(ids, centres, n_dir=8):
    """analytic direction class (1..n_dir, 0 = background) of the vector pixel -> own centre,
    quantised like DTOffsetHelper.align_angle (SegFix_offset_helper.py:311-341): upper-inclusive
    bins centred on -180 + k*360/n."""
    H, W = ids.shape
    cy = np.zeros(ids.max() + 1, dtype=np.float64)
    cx = np.zeros(ids.max() + 1, dtype=np.float64)
    cy[1:len(centres) + 1] = centres[:, 0]
    cx[1:len(centres) + 1] = centres[:, 1]
    yy, xx = np.mgrid[0:H, 0:W]
    dy = cy[ids] - yy
    dx = cx[ids] - xx
    ang = np.degrees(np.arctan2(dy, dx))  # integer inputs: never within an ulp of a bin edge
    step = 360.0 / n_dir
    idx = np.ceil((ang + 180.0 - step / 2.0) / step).astype(np.int64) % n_dir
    return np.where(ids > 0, idx + 1, 0).astype(np.uint8)


def _gauss17():
    t = np.arange(-8, 9, dtype=np.float64)
    w = np.exp(-0.5 * t * t / 4.0)
    w = w / w.sum()
    return _q(w)


def postproc_inputs(seed, H, W, n_target=None, n_dir=8, flip_frac=0.02):
    """Inputs of the direction-aware post-processing (test_dam.py:455-563) for one tile.

    Returns dict(dcm u8 [8,H,W], prob f32 [3,H,W], point f32 [1,H,W], ids int32 [H,W]).
    """
    if n_target is None:
        n_target = int(round(700.0 * H * W / 1.0e6))
    rng = np.random.default_rng(seed + 7919)
    ids, centres = instance_map(seed, H, W, n_target, return_centres=True)
    fg = ids > 0
    ring = _cross_extreme(ids, np.maximum) != _cross_extreme(ids, np.minimum)
    gt = np.where(fg, 1, 0).astype(np.int64)
    gt[ring & fg] = 2
    gt[ring & ~fg] = 2
    # structured errors: holes punched into nuclei, specks in the background
    n_holes = max(1, len(centres) // 10)
    for i in rng.integers(0, max(1, len(centres)), size=n_holes):
        if len(centres) == 0:
            break
        y, x = centres[i]
        r = int(rng.integers(1, 3))
        gt[max(0, y - r):y + r + 1, max(0, x - r):x + r + 1] = 0
    n_specks = max(1, (H * W) // 20000)
    sy = rng.integers(0, H, size=n_specks)
    sx = rng.integers(0, W, size=n_specks)
    sr = rng.integers(1, 4, size=n_specks)
    for y, x, r in zip(sy, sx, sr):
        blk = gt[max(0, y - r):y + r + 1, max(0, x - r):x + r + 1]
        blk[blk == 0] = 1
    w = np.empty((3, H, W), dtype=np.float32)
    for c in range(3):
        w[c] = np.float32(1.0) + np.float32(5.0) * (gt == c).astype(np.float32) + _irwin_hall(rng, (H, W))
    prob = w / w.sum(axis=0, keepdims=True, dtype=np.float32)
    # 8 TTA direction-argmax maps
    base = centripetal_classes(ids, centres, n_dir)
    base[gt != 1] = 0  # the network's direction argmax is gated by the background probability
    dcm = np.empty((8, H, W), dtype=np.uint8)
    for t in range(8):
        m = base.copy()
        dy, dx = int(rng.integers(-1, 2)), int(rng.integers(-1, 2))
        m = _shift_zero(m, dy, dx)
        flip = (rng.integers(0, 1 << 16, size=(H, W)) < int(flip_frac * (1 << 16))) & (m > 0)
        m[flip] = rng.integers(0, n_dir + 1, size=int(flip.sum()), dtype=np.int64).astype(np.uint8)
        dcm[t] = m
    # point map: Gaussian blobs (sigma 2, peak 255*w0^2 ~ 10.15) at the centres + clipped noise
    pt = np.zeros((H, W), dtype=np.float64)
    g = _gauss17()
    blob = 255.0 * np.outer(g, g)
    for (y, x) in centres:
        y0, y1, x0, x1 = max(0, y - 8), min(H, y + 9), max(0, x - 8), min(W, x + 9)
        pt[y0:y1, x0:x1] += blob[y0 - y + 8:y1 - y + 8, x0 - x + 8:x1 - x + 8]
    noise = _irwin_hall(rng, (H, W)) - np.float32(2.0)
    point = np.maximum(pt.astype(np.float32) + noise * np.float32(0.75), np.float32(0.0))
    return {"dcm": dcm, "prob": prob.astype(np.float32), "point": point[None].astype(np.float32),
            "ids": ids}


def postproc_edge_cases():
    """(name, dict(dcm, prob, point)) degenerate post-processing inputs: point map
    without a positive maximum, no foreground, only foreground, unknown class ids, exact probability ties,
    tiles down to 1 x 1, a single nucleus."""
    base = postproc_inputs(901, 72, 88, 8)

    def cp():
        return {k: base[k].copy() for k in ("dcm", "prob", "point")}
    c = cp(); c["point"][:] = 0
    yield "point_all_zero", c
    c = cp(); c["point"][:] = -1.5
    yield "point_all_negative", c
    c = cp(); c["point"] = -np.abs(c["point"]) - np.float32(0.25)
    yield "point_negative_varied", c
    c = cp(); c["point"][:] = 3.0
    yield "point_constant_positive", c
    c = cp(); c["prob"][0] = 5; c["prob"][1] = 0; c["prob"][2] = 0
    yield "all_background", c
    c = cp(); c["prob"][0] = 0; c["prob"][1] = 5; c["prob"][2] = 0
    yield "all_inside", c
    c = cp(); c["dcm"][:, 10:30, 10:40] = 200
    yield "unknown_class_ids", c
    c = cp(); c["prob"][:, :, :] = 0.25
    yield "prob_ties", c
    for H, W in ((1, 1), (1, 37), (41, 1), (2, 2), (3, 5), (8, 4)):
        d = postproc_inputs(902, max(H, 24), max(W, 24), 3)
        yield "tiny_%dx%d" % (H, W), {k: np.ascontiguousarray(d[k][..., :H, :W]) for k in ("dcm", "prob", "point")}
    d = postproc_inputs(903, 64, 64, 1)
    yield "single_nucleus", {k: d[k] for k in ("dcm", "prob", "point")}


def label_edge_cases(H=48, W=56):
    """(name, uint8 [H,W] id plane) degenerate inputs of the target transform: empty, one id everywhere, binary
    masks (three-class branch), nuclei below the 5-pixel limit, touching / diagonal pairs, 1- and 3-pixel lines,
    nuclei on the frame, holes, nesting, one id on two blobs, a tile without background, salt-and-pepper ids."""
    z = np.zeros((H, W), np.uint8)
    yield "empty", z.copy()
    a = z.copy(); a[:] = 9
    yield "one_id_everywhere", a
    a = z.copy(); a[10:30, 10:30] = 255
    yield "binary_255", a
    a = z.copy(); a[10:30, 10:30] = 100
    yield "binary_100", a
    a = z.copy(); a[10:12, 10:12] = 7
    yield "tiny", a
    a = z.copy(); a[10:12, 10:12] = 7; a[30:32, 30:33] = 9
    yield "two_tiny", a
    a = z.copy(); a[5:25, 5:25] = 3; a[5:25, 25:45] = 4
    yield "touching_pair", a
    a = z.copy(); a[5:25, 5:25] = 3; a[25:45, 25:45] = 4
    yield "diagonal_pair", a
    a = z.copy(); a[20, 5:50] = 3; a[30:40, 10:30] = 8
    yield "line_1px", a
    a = z.copy(); a[20:23, 5:50] = 3; a[30:40, 10:30] = 8
    yield "line_3px", a
    a = z.copy(); a[0:14, 0:17] = 3; a[30:48, 40:56] = 9; a[20:30, 20:33] = 200
    yield "frame", a
    a = z.copy(); a[8:40, 8:48] = 5; a[18:28, 20:34] = 0; a[2:6, 2:6] = 6
    yield "ring_with_hole", a
    a = z.copy(); a[8:40, 8:48] = 5; a[18:28, 20:34] = 7
    yield "nested", a
    a = z.copy(); a[4:20, 4:20] = 5; a[28:44, 30:50] = 5; a[24:27, 4:10] = 6
    yield "same_id_two_blobs", a
    yield "no_background_stripes", (np.arange(H * W).reshape(H, W) % 7 + 1).astype(np.uint8)
    rng = np.random.default_rng(1)
    yield "noise", (rng.random((H, W)) < 0.5).astype(np.uint8) * rng.integers(1, 5, (H, W)).astype(np.uint8)


def training_inputs(seed=31):
    """Seeded inputs of the training-side operators (quantiser, one-hot, label image): dict of arrays."""
    rng = np.random.default_rng(seed)
    special = [22.5, 22.500002, 67.5, -157.5, 157.5, 180.0, -180.0, 0.0, -0.0, 11.25, -168.75, 168.75,
               np.nan, np.inf, -np.inf, 1e30, -1e30]
    ang = np.concatenate([np.linspace(-180, 180, 1441), special, rng.uniform(-200, 200, 470)])
    vec = rng.standard_normal((40, 56, 2))
    vec[0, :8] = [[0, 0], [0, 1], [1, 0], [0, -1], [-1, 0], [1, 1], [-1, -1], [1e-30, -1]]
    out = {"angle64": ang.reshape(8, -1).astype(np.float64), "angle32": ang.reshape(8, -1).astype(np.float32),
           "vec64": vec.astype(np.float64), "vec32": vec.astype(np.float32)}
    B, H, W = 4, 40, 52
    ids = np.stack([instance_map(seed + 1 + i, H, W, 7) for i in range(B)])
    cls = np.stack([centripetal_classes(*instance_map(seed + 1 + i, H, W, 7, return_centres=True), n_dir=8)
                    for i in range(B)]).astype(np.int64)
    tern = np.where(ids > 0, 1, 0).astype(np.int64)
    tern[:, ::7, :] = np.where(ids[:, ::7, :] > 0, 2, 0)
    cls[2] = 5        # a tile with ONE distinct direction value (train_util_dam.py:141)
    cls[3] = 0        # an all-background tile
    out.update(onehot_dir=cls, onehot_target=tern)
    out["labels17"] = rng.integers(0, 19, size=(2, 21, 33)).astype(np.int64)
    return out


def contiguous_ids(ids):
    """ids renumbered 1..N in ascending order of the old id (what stats_utils.remap_label does), int32."""
    u = np.unique(ids)
    u = u[u != 0]
    lut = np.zeros(int(u.max()) + 1 if u.size else 1, dtype=np.int32)
    lut[u] = np.arange(1, u.size + 1, dtype=np.int32)
    return lut[ids]


def metric_pair(seed, H, W, n_target, mode=1):
    """(true, pred) int32 label images with contiguous ids for the instance metrics (stats_utils.py).
    mode 0: two unrelated instance maps; mode 1: pred = true shifted by (2, -3) px with ~1/6 of the nuclei
    dropped and ~1/10 merged into another id -- misses, false positives and splits like a real prediction."""
    true = instance_map(seed, H, W, n_target).astype(np.int32)
    if mode == 0:
        pred = instance_map(seed + 1000, H, W, n_target).astype(np.int32)
    else:
        pred = np.roll(true, (2, -3), axis=(0, 1)).copy()
        rng = np.random.default_rng(seed)
        ids = np.unique(pred)
        ids = ids[ids > 0]
        if ids.size:
            for v in rng.choice(ids, size=max(1, ids.size // 6), replace=False):
                pred[pred == v] = 0
            for v in rng.choice(ids, size=max(1, ids.size // 10), replace=False):
                pred[pred == v] = int(rng.choice(ids))
    return contiguous_ids(true), contiguous_ids(pred)
