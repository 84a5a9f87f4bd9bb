"""numpy/scipy stand-in for cdnet_b200.sharded.CudaBackend -- TEST INFRASTRUCTURE ONLY.

Lets the host logic of the whole-slide row partition (halo exchange, seam union-find, owner / id
hand-out) run on CPU under gloo.  It mirrors the contract of the cdnet_shard_* kernels: L = flat index
of the component's first pixel in the extended tile, per-root attributes stored AT the root pixel."""
import numpy as np
from scipy import ndimage as ndi

from oracle import restate as O

FULL = np.ones((3, 3), bool)


def _roots(lab):
    """label image (0 = none) -> plane of min flat index per label (garbage where lab == 0)"""
    n = int(lab.max())
    idx = np.arange(lab.size).reshape(lab.shape)
    if n == 0:
        return idx.astype(np.int32)
    mins = np.asarray(ndi.minimum(idx, lab, index=np.arange(1, n + 1))).astype(np.int64)
    table = np.concatenate([[0], mins])
    return np.where(lab > 0, table[lab], idx).astype(np.int32)


def _f32_to_ordered(f):
    u = np.float32(f).view(np.uint32)
    return int((~u) & 0xffffffff) if (u & 0x80000000) else int(u | 0x80000000)


def _ordered_to_f32(u):
    u = np.uint32(u)
    v = (u & np.uint32(0x7fffffff)) if (u & np.uint32(0x80000000)) else ~u
    return np.uint32(v).view(np.float32)


class NumpyBackend(object):
    def to_dev(self, a):
        return np.ascontiguousarray(a)

    def to_host(self, t):
        return np.asarray(t)

    def zeros(self, shape, dtype):
        return np.zeros(shape, dtype=dtype)

    def empty(self, shape, dtype):
        return np.zeros(shape, dtype=dtype)

    def scatter(self, plane, flat_idx, vals):
        if len(flat_idx):
            plane.reshape(-1)[np.asarray(flat_idx, dtype=np.int64)] = np.asarray(vals).astype(plane.dtype)

    def gather(self, plane, flat_idx):
        return plane.reshape(-1)[np.asarray(flat_idx, dtype=np.int64)]

    def add_scalar(self, plane, v):
        plane += np.int32(v)

    def to_torch(self, t):
        import torch
        return torch.from_numpy(np.ascontiguousarray(t))

    def from_torch(self, t):
        return t.numpy() if hasattr(t, "numpy") else t

    def seam_gid(self, L, valid, r0, off):
        g = L[r0:r0 + 2].astype(np.int64) + int(off)
        if valid is not None:
            g = np.where(valid[r0:r0 + 2] != 0, g, -1)
        return np.ascontiguousarray(g.astype(np.int32))

    def seam_export(self, L, valid, attr, off, He, has_top, has_bottom, nb_top):
        parts = []
        for side, r0 in (("top", 0), ("bottom", He - 2)):
            if (side == "top" and not has_top) or (side == "bottom" and not has_bottom):
                continue
            g = L[r0:r0 + 2].astype(np.int64) + int(off)
            if valid is not None:
                g = np.where(valid[r0:r0 + 2] != 0, g, -1)
            a = attr.reshape(-1)[np.maximum(g - int(off), 0)].astype(np.int64) if attr is not None else np.zeros_like(g)
            nb = np.asarray(nb_top).astype(np.int64) if (side == "bottom" and nb_top is not None) else np.full_like(g, -1)
            parts.append(np.stack([g.reshape(-1), nb.reshape(-1), a.reshape(-1)], axis=1))
        if not parts:
            return np.zeros((0, 3), np.int64)
        rows = np.concatenate(parts, axis=0)
        return rows[rows[:, 0] >= 0]

    def ddm_codes(self, dcm_ext, n_classes, row_lo, row_hi):
        T = dcm_ext.shape[0]
        codes = np.zeros(dcm_ext.shape[1:], np.uint16)
        flags = 0
        for t in range(T):
            d = O.ddm_codes(dcm_ext[t], n_classes).astype(np.uint16)
            codes |= d << (2 * t)
            for v in np.unique(d[row_lo:row_hi]):
                flags |= 1 << (3 * t + int(v))
        return codes, flags

    def point_max(self, point_own):
        return _f32_to_ordered(np.max(point_own))

    def boost(self, codes, flags, point_ext, pmax, prob_ext, n_maps):
        vals = []
        for t in range(n_maps):
            f = (flags >> (3 * t)) & 7
            present = [v for v in range(3) if f & (1 << v)]
            mn, mx = min(present), max(present)
            d = ((codes >> (2 * t)) & 3).astype(np.float32)
            vals.append((d - np.float32(mn)) / np.float32(mx - mn))
        ddm = vals[0] if n_maps == 1 else np.mean(np.stack(vals, axis=2).astype(np.float64), axis=2)
        mxv = _ordered_to_f32(pmax)
        with np.errstate(invalid="ignore", divide="ignore"):
            gate = (point_ext.astype(np.float32) / mxv > 0.2) * 1
        gate = O.dilate(gate, O.disk(1))
        eb = 2 * (ddm - ddm * gate)
        p = prob_ext.astype(np.float32).copy()
        p[2] = (p[2] + 0.5 * eb) * (1 + eb)
        return (np.argmax(p, axis=0) == 1).astype(np.uint8)

    def stage1(self, inside, top_frame, bottom_frame):
        fg = inside != 0
        lf = ndi.label(fg)[0]
        lb = ndi.label(~fg)[0]
        lab = np.where(fg, lf, lb + lf.max())
        L = _roots(lab)
        frame = np.zeros(fg.shape, bool)
        frame[:, 0] = frame[:, -1] = True
        if top_frame:
            frame[0] = True
        if bottom_frame:
            frame[-1] = True
        touch = np.zeros(fg.shape, np.int32)
        touch.reshape(-1)[np.unique(L[frame & ~fg])] = 1
        return L, touch

    def stage2(self, inside, L, touch, row_lo, row_hi):
        fg = inside != 0
        state = np.where(fg, 1, np.where(touch.reshape(-1)[L] != 0, 0, 2)).astype(np.uint8)
        lab = ndi.label(state != 0)[0]
        L[...] = np.where(state != 0, _roots(lab), L)
        area = np.zeros(fg.shape, np.int32)
        own = np.zeros(fg.shape, bool)
        own[row_lo:row_hi] = True
        r, c = np.unique(L[(state != 0) & own], return_counts=True)
        area.reshape(-1)[r] = c
        return state, area

    def stage3(self, state, L, area, min_area):
        keep = ((state != 0) & (area.reshape(-1)[L] >= min_area)).astype(np.uint8)
        lab = ndi.label(keep != 0, FULL)[0]
        L[...] = np.where(keep != 0, _roots(lab), L)
        return keep

    def stage4(self, L, keep, excluded):
        idx = np.arange(L.size).reshape(L.shape)
        roots = (keep != 0) & (L == idx) & (excluded == 0)
        idmap = np.zeros(L.shape, np.int32)
        n = int(roots.sum())
        idmap[roots] = np.arange(1, n + 1)
        return idmap, n

    def relabel(self, L, keep, idmap, out=None):
        res = np.where(keep != 0, idmap.reshape(-1)[L], 0).astype(np.int32)
        if out is None:
            return res
        out[...] = res
        return out

    def dilate(self, labels_ext, radius, out_dtype):
        return O.dilate(labels_ext, O.disk(radius)).astype(out_dtype)
