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

    def stack(self, arrays):
        return np.ascontiguousarray(np.stack(arrays, axis=0))

    def to_torch(self, t):
        import torch
        return torch.from_numpy(np.ascontiguousarray(t))

    def from_torch(self, t):
        return t.numpy() if hasattr(t, "numpy") else t

    # -- scalars
    def pack_scalars(self, flags, pmax):
        return np.array([flags[0], pmax[0]], dtype=np.int32)

    def combine_scalars(self, gathered):
        g = np.asarray(gathered)
        flags = np.bitwise_or.reduce(g[:, 0].astype(np.int64) & 0xffffffff)
        pm = (g[:, 1].astype(np.int64) & 0xffffffff).max()
        return (np.array([flags], dtype=np.int64).astype(np.uint32).view(np.int32),
                np.array([pm], dtype=np.int64).astype(np.uint32).view(np.int32))

    def flags_to_host(self, flags):
        return int(np.asarray(flags).view(np.uint32)[0])

    def exclusive_offset(self, gathered_counts, rank):
        return np.array([int(np.asarray(gathered_counts).reshape(-1)[:rank].sum())], dtype=np.int32)

    def add_offset(self, plane, offset):
        plane += np.int32(offset[0])

    # -- seam rounds: same table contract as csrc/seam.cu, solved with scipy on the host
    def seam_init(self, sh, world, W):
        sh.cap = 4 * W + 1
        sh.emitted = np.zeros((sh.He, W), np.int32)
        sh.world = world

    def seam_gid(self, L, valid, r0, off):
        g = L[r0:r0 + 2].astype(np.int64) + int(off)
        if valid is not None:
            g = np.where(valid[r0:r0 + 2] != 0, g, -1)
        return np.ascontiguousarray(g.astype(np.int32))

    def seam_export(self, sh, valid, attr, round_id, nb_gid):
        W = sh.L.shape[1]
        rows = []
        err = 0
        for side, r0 in ((0, 0), (1, sh.He - 2)):
            if (side == 0 and not sh.has_top) or (side == 1 and not sh.has_bottom):
                continue
            g = self.seam_gid(sh.L, valid, r0, sh.off).astype(np.int64)
            nb = np.asarray(nb_gid).astype(np.int64) if (side == 1 and nb_gid is not None) else np.full_like(g, -1)
            if side == 1 and nb_gid is not None and ((g >= 0) != (nb >= 0)).any():
                err |= 1
            for r in range(2):
                start = np.ones(W, bool)
                start[1:] = (g[r, 1:] != g[r, :-1]) | (nb[r, 1:] != nb[r, :-1])
                sel = start & (g[r] >= 0)
                for x in np.nonzero(sel)[0]:
                    root = int(g[r, x] - sh.off)
                    a = int(attr.reshape(-1)[root]) if attr is not None else 0
                    first = sh.emitted.reshape(-1)[root] != round_id
                    sh.emitted.reshape(-1)[root] = round_id
                    rows.append((int(g[r, x]), int(nb[r, x]), a, a if first else 0))
        tbl = np.zeros((sh.cap, 4), np.int32)
        tbl[0] = (len(rows), err, 0, 0)
        if rows:
            tbl[1:1 + len(rows)] = np.asarray(rows, dtype=np.int64).astype(np.int32)
        return tbl

    @staticmethod
    def _classes(gathered):
        from scipy.sparse import coo_matrix
        from scipy.sparse.csgraph import connected_components
        ent = [np.asarray(gathered[r][1:1 + int(gathered[r][0, 0])]).astype(np.int64) for r in range(len(gathered))]
        allent = np.concatenate(ent) if ent else np.zeros((0, 4), np.int64)
        keys = np.unique(np.concatenate([allent[:, 0], allent[:, 1][allent[:, 1] >= 0]])) if len(allent) else np.zeros(0, np.int64)
        if keys.size == 0:
            return ent, keys, np.zeros(0, np.int64), 0
        e = allent[allent[:, 1] >= 0]
        a, b = np.searchsorted(keys, e[:, 0]), np.searchsorted(keys, e[:, 1])
        ncls, cls = connected_components(coo_matrix((np.ones(a.size, np.int8), (a, b)), shape=(keys.size, keys.size)),
                                         directed=False)
        return ent, keys, cls.astype(np.int64), int(ncls)

    def seam_solve(self, sh, gathered, mode, W, plane=None, excluded=None):
        ent, keys, cls, ncls = self._classes(gathered)
        sh._seam = (ent, keys, cls, ncls)
        if ncls == 0:
            return
        allent = np.concatenate(ent)
        cidx = cls[np.searchsorted(keys, allent[:, 0])]
        mine = ent[sh.rank]
        mcls = cls[np.searchsorted(keys, mine[:, 0])] if len(mine) else np.zeros(0, np.int64)
        if mode == 0:
            val = np.zeros(ncls, np.int64)
            np.maximum.at(val, cidx, (allent[:, 2] != 0).astype(np.int64))
            plane.reshape(-1)[mine[:, 0] - sh.off] = val[mcls].astype(np.int32)
        elif mode == 1:
            val = np.zeros(ncls, np.int64)
            np.add.at(val, cidx, allent[:, 3])
            plane.reshape(-1)[mine[:, 0] - sh.off] = np.minimum(val[mcls], 2 ** 31 - 1).astype(np.int32)
        else:
            croot = np.full(ncls, np.iinfo(np.int64).max, np.int64)
            np.minimum.at(croot, cls, keys)
            sh._croot = croot
            g = mine[:, 0]
            ex = (croot[mcls] != g) | (g < sh.r0 * W) | (g >= sh.r1 * W)
            excluded.reshape(-1)[(g - sh.off)[ex]] = 1

    def seam_ids_export(self, sh, gathered, idmap, round_id, W):
        ent, keys, cls, ncls = sh._seam
        mine = ent[sh.rank]
        tbl2 = np.zeros((sh.cap, 4), np.int32)
        if len(mine):
            g = np.unique(mine[:, 0])
            own = (sh._croot[cls[np.searchsorted(keys, g)]] == g) & (g >= sh.r0 * W) & (g < sh.r1 * W)
            go = g[own]
            tbl2[0, 0] = len(go)
            tbl2[1:1 + len(go), 0] = go.astype(np.int32)
            tbl2[1:1 + len(go), 1] = idmap.reshape(-1)[go - sh.off]
        return tbl2

    def seam_ids_apply(self, sh, gathered, gathered2, idmap):
        ent, keys, cls, ncls = sh._seam
        mine = ent[sh.rank]
        if not len(mine):
            return
        t2 = np.concatenate([np.asarray(gathered2[r][1:1 + int(gathered2[r][0, 0])]).astype(np.int64)
                             for r in range(len(gathered2))])
        order = np.argsort(t2[:, 0])
        tk, tv = t2[order, 0], t2[order, 1]
        groot = sh._croot[cls[np.searchsorted(keys, mine[:, 0])]]
        pos = np.searchsorted(tk, groot)
        assert np.array_equal(tk[pos], groot), "every seam class must have exactly one owner"
        idmap.reshape(-1)[mine[:, 0] - sh.off] = tv[pos].astype(np.int32)

    def seam_errors(self, tables):
        return int(max(int(t[0, 1]) for t in tables)) if tables else 0

    def ddm_codes(self, dcm_ext, n_classes, row_lo, row_hi):
        T = dcm_ext.shape[0]
        codes = np.zeros(dcm_ext.shape[1:], np.uint16)
        flags = 0
        for t in range(T):
            d = O.ddm_codes(dcm_ext[t], n_classes).astype(np.uint16)
            codes |= d << (2 * t)
            for v in np.unique(d[row_lo:row_hi]):
                flags |= 1 << (3 * t + int(v))
        return codes, np.array([flags], dtype=np.uint32).view(np.int32)

    def point_max(self, point_own):
        return np.array([_f32_to_ordered(np.max(point_own))], dtype=np.uint32).view(np.int32)

    def boost(self, codes, flags, point_ext, pmax, prob_ext, n_maps):
        flags = int(np.asarray(flags).view(np.uint32)[0])
        pmax = int(np.asarray(pmax).view(np.uint32)[0])
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
        return idmap, np.array([n], dtype=np.int32)

    def relabel(self, L, keep, idmap, out=None):
        res = np.where(keep != 0, idmap.reshape(-1)[L], 0).astype(np.int32)
        if out is None:
            return res
        out[...] = res
        return out

    def dilate(self, labels_ext, radius, out_dtype):
        return O.dilate(labels_ext, O.disk(radius)).astype(out_dtype)
