"""Multi-GPU execution of the geometry hot path.

* Tile batches (BASELINE configs 2-4) are independent units: `shard_tiles` hands every rank its
  contiguous block, there is NO data-path collective (bench.py --gpus N).
* A whole-slide prediction map (config 5) is ROW-PARTITIONED: every rank post-processes its block of
  rows (test_dam.py:455-563 semantics for the whole slide, postproc = 0) and the ranks exchange only
    - halo rows (1 row of class map / point map / inside mask, 2 rows of labels),
    - two scalars (DDM value-present flags, point-map maximum),
    - the roots of the components that touch a shard seam: a small union-find over those roots, solved
      redundantly on every rank's host, makes fill-holes (a hole is a background component that never
      reaches the SLIDE frame), remove-small (areas summed over shards) and the canonical raster-order
      numbering (ids = rank of the component's first pixel in the whole slide) identical to a
      single-GPU run.
  The per-rank compute is the same sm_100a kernels, cut into stages (cdnet_shard_* in
  include/cdnet_b200.h); the host logic below is backend-agnostic so that it can be exercised on CPU
  with gloo (tests/test_sharded_gloo.py provides a numpy stand-in backend; the product has only the
  CUDA one).

`comm` abstracts who is local: DistComm = one rank per process (torch.distributed, NCCL on the GPU
box / gloo in the CPU tests), SimComm = all ranks in this process (used to check sharded == unsharded
on a single GPU).
"""
import numpy as np


# --------------------------------------------------------------------------------------------------
# partitioning
# --------------------------------------------------------------------------------------------------
def shard_tiles(n_tiles, world, rank):
    """contiguous block of tile indices of `rank` (sizes differ by at most one)"""
    base, rem = divmod(n_tiles, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def row_partition(H, world):
    """[(r0, r1)] contiguous row blocks; every block has at least 2 rows"""
    assert H >= 2 * world, "a shard needs at least 2 rows"
    return [shard_tiles(H, world, r) for r in range(world)]


# --------------------------------------------------------------------------------------------------
# communicators
# --------------------------------------------------------------------------------------------------
class SimComm(object):
    """all `world` ranks live in this process (lists are already global)"""

    def __init__(self, world):
        self.world = world
        self.local_ranks = list(range(world))

    def allgather(self, values):
        assert len(values) == self.world
        return list(values)


class DistComm(object):
    """one rank per process over torch.distributed (NCCL with CUDA tensors, gloo on CPU)"""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.local_ranks = [self.rank]

    def allgather(self, values):
        """values: [one picklable object] -> list over all ranks"""
        assert len(values) == 1
        out = [None] * self.world
        self.dist.all_gather_object(out, values[0], group=self.group)
        return out


# --------------------------------------------------------------------------------------------------
# seam reconciliation (host logic; numpy + scipy.sparse.csgraph)
# --------------------------------------------------------------------------------------------------
def seam_classes(seam_info, world):
    """Union of the rank-local components that meet at shard seams.

    seam_info[r] = dict(top=(gid[2,W], valid[2,W]) or None, bottom=(gid, valid) or None) where gid is
    the slide-global pixel index of the LOCAL root of each pixel of the two rows shared with the
    neighbour (rank r: bottom rows = own last row + ghost row; rank r+1: top rows = ghost row + own
    first row -- the same two slide rows).  Returns (keys sorted unique int64, class_of_key,
    n_classes): two keys are in one class iff their local components are connected across seams."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    a_list, b_list, all_keys = [], [], []
    for r in range(world - 1):
        lo, hi = seam_info[r]["bottom"], seam_info[r + 1]["top"]
        ga, va = lo
        gb, vb = hi
        both = va & vb
        assert np.array_equal(va, vb), "seam pixels must be classified identically on both ranks"
        a_list.append(ga[both].astype(np.int64))
        b_list.append(gb[both].astype(np.int64))
    for r in range(world):
        for side in ("top", "bottom"):
            if seam_info[r][side] is not None:
                g, v = seam_info[r][side]
                all_keys.append(g[v].astype(np.int64))
    keys = np.unique(np.concatenate(all_keys)) if all_keys else np.zeros(0, np.int64)
    if keys.size == 0:
        return keys, np.zeros(0, np.int64), 0
    a = np.searchsorted(keys, np.concatenate(a_list)) if a_list else np.zeros(0, np.int64)
    b = np.searchsorted(keys, np.concatenate(b_list)) if b_list else np.zeros(0, np.int64)
    n = keys.size
    graph = coo_matrix((np.ones(a.size, np.int8), (a, b)), shape=(n, n))
    ncls, cls = connected_components(graph, directed=False)
    return keys, cls.astype(np.int64), int(ncls)


def _class_reduce(keys, cls, ncls, entry_keys, entry_vals, how):
    """aggregate per-key values over classes; entries with duplicate keys must already be de-duplicated
    by the caller where that matters (sum)"""
    idx = cls[np.searchsorted(keys, entry_keys)]
    if how == "or":
        out = np.zeros(ncls, np.int64)
        np.maximum.at(out, idx, (entry_vals != 0).astype(np.int64))
    elif how == "sum":
        out = np.zeros(ncls, np.int64)
        np.add.at(out, idx, entry_vals.astype(np.int64))
    elif how == "min":
        out = np.full(ncls, np.iinfo(np.int64).max, np.int64)
        np.minimum.at(out, idx, entry_vals.astype(np.int64))
    else:
        raise ValueError(how)
    return out


# --------------------------------------------------------------------------------------------------
# the CUDA backend (the only one the product ships)
# --------------------------------------------------------------------------------------------------
class CudaBackend(object):
    """per-rank device ops on one extended tile; arrays are torch CUDA tensors"""

    def __init__(self, device=None):
        import torch
        from . import _cabi, api
        self.torch, self.L, self.api = torch, _cabi.lib(), api
        self.dev = api._device(device)

    # -- plumbing
    def to_dev(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a)).to(self.dev)

    def to_host(self, t):
        return t.cpu().numpy()

    def zeros(self, shape, dtype):
        return self.torch.zeros(shape, dtype=getattr(self.torch, dtype), device=self.dev)

    def empty(self, shape, dtype):
        return self.torch.empty(shape, dtype=getattr(self.torch, dtype), device=self.dev)

    def cat_rows(self, parts):
        return self.torch.cat(parts, dim=-2).contiguous()

    def scatter(self, plane, flat_idx, vals):
        if len(flat_idx):
            idx = self.torch.from_numpy(np.asarray(flat_idx, dtype=np.int64)).to(self.dev)
            v = self.torch.from_numpy(np.asarray(vals)).to(self.dev).to(plane.dtype)
            plane.view(-1)[idx] = v

    def gather(self, plane, flat_idx):
        idx = self.torch.from_numpy(np.asarray(flat_idx, dtype=np.int64)).to(self.dev)
        return plane.view(-1)[idx].cpu().numpy()

    def add_scalar(self, plane, v):
        plane += int(v)

    def _st(self):
        return self.torch.cuda.current_stream().cuda_stream

    # -- kernels
    def ddm_codes(self, dcm_ext, n_classes, row_lo, row_hi):
        from ._cabi import check
        T, He, W = dcm_ext.shape
        codes = self.empty((He, W), "uint16") if hasattr(self.torch, "uint16") else None
        if codes is None:
            codes = self.empty((He, W), "int16")
        flags = self.zeros((1,), "int32")
        check(self.L.cdnet_shard_ddm_codes(dcm_ext.data_ptr(), codes.data_ptr(), flags.data_ptr(), T, He, W,
                                           int(n_classes), int(row_lo), int(row_hi), self._st()), "shard_ddm_codes")
        return codes, int(flags.cpu().numpy().view(np.uint32)[0])

    def point_max(self, point_own):
        from ._cabi import check
        pm = self.zeros((1,), "int32")
        p = point_own.contiguous()
        check(self.L.cdnet_shard_point_max(p.data_ptr(), pm.data_ptr(), p.numel(), self._st()), "shard_point_max")
        return int(pm.cpu().numpy().view(np.uint32)[0])

    def boost(self, codes, flags, point_ext, pmax, prob_ext, n_maps):
        from ._cabi import check
        He, W = codes.shape
        inside = self.empty((He, W), "uint8")
        status = self.zeros((1,), "int32")
        f = self.to_dev(np.array([flags], dtype=np.uint32).view(np.int32))
        pm = self.to_dev(np.array([pmax], dtype=np.uint32).view(np.int32))
        check(self.L.cdnet_shard_boost(codes.data_ptr(), f.data_ptr(), point_ext.data_ptr(), pm.data_ptr(),
                                       prob_ext.data_ptr(), inside.data_ptr(), status.data_ptr(), He, W, int(n_maps), 0,
                                       self._st()), "shard_boost")
        return inside

    def stage1(self, inside, top_frame, bottom_frame):
        from ._cabi import check
        He, W = inside.shape
        L = self.empty((He, W), "int32")
        touch = self.empty((He, W), "int32")
        check(self.L.cdnet_shard_label_stage1(inside.data_ptr(), L.data_ptr(), touch.data_ptr(), He, W,
                                              1 if top_frame else 0, 1 if bottom_frame else 0, self._st()), "stage1")
        return L, touch

    def stage2(self, inside, L, touch, row_lo, row_hi):
        from ._cabi import check
        He, W = inside.shape
        state = self.empty((He, W), "uint8")
        area = self.zeros((He, W), "int32")
        check(self.L.cdnet_shard_label_stage2(inside.data_ptr(), L.data_ptr(), touch.data_ptr(), state.data_ptr(),
                                              area.data_ptr(), He, W, int(row_lo), int(row_hi), self._st()), "stage2")
        return state, area

    def stage3(self, state, L, area, min_area):
        from ._cabi import check
        He, W = state.shape
        keep = self.empty((He, W), "uint8")
        check(self.L.cdnet_shard_label_stage3(state.data_ptr(), L.data_ptr(), area.data_ptr(), keep.data_ptr(),
                                              int(min_area), He, W, self._st()), "stage3")
        return keep

    def stage4(self, L, keep, excluded):
        from ._cabi import check
        He, W = keep.shape
        idmap = self.empty((He, W), "int32")
        rowcnt = self.empty((He,), "int32")
        n = self.zeros((1,), "int32")
        check(self.L.cdnet_shard_label_stage4(L.data_ptr(), keep.data_ptr(), excluded.data_ptr(), idmap.data_ptr(),
                                              rowcnt.data_ptr(), n.data_ptr(), He, W, self._st()), "stage4")
        return idmap, int(n.cpu().numpy()[0])

    def relabel(self, L, keep, idmap):
        from ._cabi import check
        He, W = keep.shape
        labels = self.empty((He, W), "int32")
        check(self.L.cdnet_shard_relabel(L.data_ptr(), keep.data_ptr(), idmap.data_ptr(), labels.data_ptr(), He, W,
                                         self._st()), "relabel")
        return labels

    def dilate(self, labels_ext, radius, out_dtype):
        t = self.api.label_dilate_cuda(labels_ext[None], radius,
                                       out_dtype=getattr(self.torch, out_dtype))
        return t[0]


# --------------------------------------------------------------------------------------------------
# whole-slide post-processing
# --------------------------------------------------------------------------------------------------
class _Shard(object):
    pass


def _seam_rows(sh):
    """ext-row indices of the two rows shared with the upper / lower neighbour"""
    top = (0, 1) if sh.has_top else None
    bottom = (sh.He - 2, sh.He - 1) if sh.has_bottom else None
    return top, bottom


def _seam_info(be, sh, L, valid_plane):
    """per side: (gid[2,W] int64, valid[2,W] bool) of the rows shared with the neighbour"""
    info = {"top": None, "bottom": None}
    top, bottom = _seam_rows(sh)
    for side, rows in (("top", top), ("bottom", bottom)):
        if rows is None:
            continue
        r0 = rows[0]
        g = be.to_host(L[r0:r0 + 2]).astype(np.int64) + sh.off
        v = be.to_host(valid_plane[r0:r0 + 2]) != 0 if valid_plane is not None else np.ones(g.shape, bool)
        info[side] = (g, v)
    return info


def _seam_entries(be, sh, info, attr_plane):
    """(keys, vals) of this rank's seam roots, de-duplicated per local root: attr_plane[root]"""
    ks = [info[s][0][info[s][1]] for s in ("top", "bottom") if info[s] is not None]
    if not ks:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    keys = np.unique(np.concatenate(ks))
    vals = be.gather(attr_plane, keys - sh.off) if attr_plane is not None else np.zeros(keys.size, np.int64)
    return keys, np.asarray(vals).astype(np.int64)


def postprocess_slide(shards, comm, H, W, be, direction_classes=9, min_area=20, radius=2, out_dtype="int64"):
    """Direction-aware post-processing (test_dam.py:455-563, postproc = 0) of an H x W slide whose rows
    are partitioned over comm.world ranks.

    shards: one dict per LOCAL rank (comm.local_ranks order) with the rank's OWN rows as numpy or
    backend arrays: dcm uint8 [T,Hl,W] (T = 1 or 8), prob float32 [3,Hl,W], point float32 [1,Hl,W].
    Returns the list of label arrays [Hl,W] (backend arrays) of the local ranks.  Raises the
    reference's AssertionError for a constant direction map."""
    G = comm.world
    parts = row_partition(H, G)
    S = []
    for rank, d in zip(comm.local_ranks, shards):
        sh = _Shard()
        sh.rank, (sh.r0, sh.r1) = rank, parts[rank]
        sh.has_top, sh.has_bottom = rank > 0, rank < G - 1
        sh.Hl = sh.r1 - sh.r0
        sh.lo = 1 if sh.has_top else 0            # ext row of the first own row
        sh.He = sh.Hl + sh.lo + (1 if sh.has_bottom else 0)
        sh.off = (sh.r0 - sh.lo) * W              # slide-global index of ext pixel 0
        asdev = lambda a: a if not isinstance(a, np.ndarray) else be.to_dev(a)
        sh.dcm, sh.prob, sh.point = asdev(d["dcm"]), asdev(d["prob"]), asdev(d["point"])
        assert sh.dcm.shape[-2] == sh.Hl and sh.dcm.shape[-1] == W
        S.append(sh)
    n_maps = int(S[0].dcm.shape[0])

    def halo(get_rows, nrows):
        """exchange the first / last `nrows` own rows with the neighbours (all-gather of the seam rows)"""
        mine = [(be.to_host(get_rows(sh, 0, nrows)), be.to_host(get_rows(sh, sh.Hl - nrows, sh.Hl))) for sh in S]
        allr = comm.allgather(mine)
        out = []
        for sh in S:
            above = be.to_dev(allr[sh.rank - 1][1]) if sh.has_top else None
            below = be.to_dev(allr[sh.rank + 1][0]) if sh.has_bottom else None
            out.append((above, below))
        return out

    def extend(sh, own, above, below):
        parts_ = ([above] if above is not None else []) + [own] + ([below] if below is not None else [])
        return be.cat_rows(parts_) if len(parts_) > 1 else own.contiguous()

    # ---- phase 1: DDM codes (1-row class-map halo), global value flags and point maximum
    h_dcm = halo(lambda sh, a, b: sh.dcm[:, a:b], 1)
    h_pt = halo(lambda sh, a, b: sh.point[:, a:b], 1)
    loc = []
    for sh, (da, db), (pa, pb) in zip(S, h_dcm, h_pt):
        sh.dcm_ext = extend(sh, sh.dcm, da, db)
        sh.point_ext = extend(sh, sh.point, pa, pb)
        sh.codes, fl = be.ddm_codes(sh.dcm_ext, direction_classes, sh.lo, sh.lo + sh.Hl)
        loc.append((fl, be.point_max(sh.point)))
    allv = comm.allgather(loc)
    flags = 0
    for f, _ in allv:
        flags |= int(f)
    pmax = max(int(p) for _, p in allv)
    for t in range(n_maps):
        f = (flags >> (3 * t)) & 7
        if f in (0, 1, 2, 4):
            raise AssertionError("constant direction map: generate_dd_map is NaN (test_dam.py:535)")

    # ---- phase 2: boost + argmax on own rows, then 1-row halo of the inside mask
    for sh in S:
        if sh.has_top or sh.has_bottom:
            # prob needs no halo (pointwise); pad the ghost rows with a copy so that the planes line up
            pa = sh.prob[:, :1] if sh.has_top else None
            pb = sh.prob[:, -1:] if sh.has_bottom else None
            prob_ext = extend(sh, sh.prob, pa, pb)
        else:
            prob_ext = sh.prob.contiguous()
        ins = be.boost(sh.codes, flags, sh.point_ext.reshape(sh.He, W), pmax, prob_ext, n_maps)
        sh.inside_own = ins[sh.lo:sh.lo + sh.Hl]
    h_in = halo(lambda sh, a, b: sh.inside_own[a:b], 1)
    for sh, (ia, ib) in zip(S, h_in):
        sh.inside = extend(sh, sh.inside_own, ia, ib)

    # ---- phase 3: forest of equal-value components; slide-global frame-touch flags for seam components
    infos, entries = [], []
    for sh in S:
        sh.L, sh.touch = be.stage1(sh.inside, sh.rank == 0, sh.rank == G - 1)
        info = _seam_info(be, sh, sh.L, None)
        infos.append(info)
        entries.append(_seam_entries(be, sh, info, sh.touch))
    g_info, g_ent = comm.allgather(infos), comm.allgather(entries)
    keys, cls, ncls = seam_classes(g_info, G)
    if ncls:
        ek = np.concatenate([e[0] for e in g_ent])
        ev = np.concatenate([e[1] for e in g_ent])
        ctouch = _class_reduce(keys, cls, ncls, ek, ev, "or")
        for sh, (k, _) in zip(S, entries):
            be.scatter(sh.touch, k - sh.off, ctouch[cls[np.searchsorted(keys, k)]].astype(np.int32))

    # ---- phase 4: fill holes, local areas of own rows; slide-global areas for seam components
    infos, entries = [], []
    for sh in S:
        sh.state, sh.area = be.stage2(sh.inside, sh.L, sh.touch, sh.lo, sh.lo + sh.Hl)
        info = _seam_info(be, sh, sh.L, sh.state)
        infos.append(info)
        entries.append(_seam_entries(be, sh, info, sh.area))
    g_info, g_ent = comm.allgather(infos), comm.allgather(entries)
    keys, cls, ncls = seam_classes(g_info, G)
    if ncls:
        # a (rank, local root) pair contributes once; equal keys on two ranks are different local parts
        ek = np.concatenate([e[0] for e in g_ent])
        ev = np.concatenate([e[1] for e in g_ent])
        carea = _class_reduce(keys, cls, ncls, ek, ev, "sum")
        for sh, (k, _) in zip(S, entries):
            be.scatter(sh.area, k - sh.off, np.minimum(carea[cls[np.searchsorted(keys, k)]], 2 ** 31 - 1).astype(np.int32))

    # ---- phase 5: remove small, 8-connectivity; owners and excluded roots; numbering
    infos, entries = [], []
    for sh in S:
        sh.keep = be.stage3(sh.state, sh.L, sh.area, min_area)
        info = _seam_info(be, sh, sh.L, sh.keep)
        infos.append(info)
        entries.append(_seam_entries(be, sh, info, None))
    g_info = comm.allgather(infos)
    keys, cls, ncls = seam_classes(g_info, G)
    croot = _class_reduce(keys, cls, ncls, keys, keys, "min") if ncls else np.zeros(0, np.int64)
    counts = []
    for sh, (k, _) in zip(S, entries):
        own_lo, own_hi = sh.r0 * W, sh.r1 * W
        groot = croot[cls[np.searchsorted(keys, k)]] if k.size else k
        excl = (groot != k) | (k < own_lo) | (k >= own_hi)
        sh.seam_keys, sh.seam_groot = k, groot
        sh.excluded = be.zeros((sh.He, W), "uint8")
        be.scatter(sh.excluded, (k - sh.off)[excl], np.ones(int(excl.sum()), np.uint8))
        # ghost-row roots that never reach an own row are not seam keys of a *kept own* pixel only if they are
        # not kept at all; kept ones appear in the shared rows and are covered above
        sh.idmap, n_owned = be.stage4(sh.L, sh.keep, sh.excluded)
        counts.append(n_owned)
    g_counts = comm.allgather(counts)
    offsets = np.concatenate([[0], np.cumsum(g_counts)])
    tables = []
    for sh in S:
        be.add_scalar(sh.idmap, offsets[sh.rank])
        k, groot = sh.seam_keys, sh.seam_groot
        own = (groot == k) & (k >= sh.r0 * W) & (k < sh.r1 * W) if k.size else np.zeros(0, bool)
        ids = be.gather(sh.idmap, (k - sh.off)[own]) if own.any() else np.zeros(0, np.int64)
        tables.append((k[own], np.asarray(ids).astype(np.int64)))
    g_tab = comm.allgather(tables)
    tk = np.concatenate([t[0] for t in g_tab]) if g_tab else np.zeros(0, np.int64)
    tv = np.concatenate([t[1] for t in g_tab]) if g_tab else np.zeros(0, np.int64)
    order = np.argsort(tk)
    tk, tv = tk[order], tv[order]
    for sh in S:
        k, groot = sh.seam_keys, sh.seam_groot
        if k.size:
            pos = np.searchsorted(tk, groot)
            assert np.array_equal(tk[pos], groot), "every seam class must have exactly one owner"
            be.scatter(sh.idmap, k - sh.off, tv[pos].astype(np.int32))
        sh.labels = be.relabel(sh.L, sh.keep, sh.idmap)

    # ---- phase 6: label dilation by disk(radius) with a `radius`-row label halo
    r = int(radius)
    outs = []
    if r > 0:
        own_labels = lambda sh, a, b: sh.labels[sh.lo:sh.lo + sh.Hl][a:b]
        h_lab = halo(own_labels, r)
    for i, sh in enumerate(S):
        own = sh.labels[sh.lo:sh.lo + sh.Hl]
        if r > 0:
            la, lb = h_lab[i]
            ext = extend(sh, own, la, lb)
            top = r if sh.has_top else 0
            out = be.dilate(ext, r, out_dtype)[top:top + sh.Hl]
        else:
            out = be.dilate(own.contiguous(), 0, out_dtype)
        outs.append(out)
    return outs
