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

    def exchange(self, be, up, down):
        """up[i] / down[i]: backend array that local rank i sends to rank-1 / rank+1 (None at the ends).
        Returns (from_upper, from_lower) lists."""
        n = self.world
        from_upper = [down[i - 1] if i > 0 else None for i in range(n)]
        from_lower = [up[i + 1] if i < n - 1 else None for i in range(n)]
        return from_upper, from_lower

    def send_up(self, be, tensors, like):
        """local rank i > 0 sends tensors[i] to rank i-1; returns what each rank receives from rank i+1"""
        n = self.world
        return [tensors[i + 1] if i < n - 1 else None for i in range(n)]


class DistComm(object):
    """one rank per process over torch.distributed (NCCL with CUDA tensors, gloo on CPU)"""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.local_ranks = [self.rank]
        # small host-side tables (seam roots, counts, id tables): all_gather_object on the default (NCCL) group
        # measured slightly faster at 8 ranks than a gloo side group (CDNET_SHARD_HOSTGROUP=gloo selects that)
        self.host_group = group
        import os
        if group is None and dist.get_backend() == "nccl" and os.environ.get("CDNET_SHARD_HOSTGROUP", "nccl") == "gloo":
            self.host_group = dist.new_group(backend="gloo")

    def allgather(self, values):
        """values: [one small picklable object] -> list over all ranks"""
        assert len(values) == 1
        out = [None] * self.world
        self.dist.all_gather_object(out, values[0], group=self.host_group)
        return out

    def exchange(self, be, up, down):
        """point-to-point halo / seam exchange with the two row neighbours (NCCL send/recv on device
        tensors, gloo on CPU tensors); what arrives from a neighbour has the shape of what is sent to it"""
        dist = self.dist
        r, n = self.rank, self.world
        ops, recv_up, recv_down = [], None, None
        t_up = be.to_torch(up[0]) if up[0] is not None else None
        t_down = be.to_torch(down[0]) if down[0] is not None else None
        if r > 0 and t_up is not None:
            recv_up = t_up.new_empty(t_up.shape)
            ops += [dist.P2POp(dist.isend, t_up, r - 1, self.group), dist.P2POp(dist.irecv, recv_up, r - 1, self.group)]
        if r < n - 1 and t_down is not None:
            recv_down = t_down.new_empty(t_down.shape)
            ops += [dist.P2POp(dist.isend, t_down, r + 1, self.group), dist.P2POp(dist.irecv, recv_down, r + 1, self.group)]
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return ([be.from_torch(recv_up) if recv_up is not None else None],
                [be.from_torch(recv_down) if recv_down is not None else None])

    def send_up(self, be, tensors, like):
        """rank r > 0 sends tensors[0] to rank r-1; rank r < n-1 receives a tensor shaped `like` from r+1"""
        dist = self.dist
        r, n = self.rank, self.world
        ops, recv = [], None
        if r > 0:
            ops.append(dist.P2POp(dist.isend, be.to_torch(tensors[0]), r - 1, self.group))
        if r < n - 1:
            recv = be.to_torch(like).new_empty(like.shape)
            ops.append(dist.P2POp(dist.irecv, recv, r + 1, self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return [be.from_torch(recv) if recv is not None else None]


# --------------------------------------------------------------------------------------------------
# seam reconciliation (host logic; numpy + scipy.sparse.csgraph)
# --------------------------------------------------------------------------------------------------
def seam_classes(edges, keys_list):
    """Union of the rank-local components that meet at shard seams.

    edges: int64 [n,2] pairs (gid_a, gid_b) -- slide-global pixel index of the LOCAL root, on the upper and
    on the lower rank, of one pixel of the two rows the ranks share; keys_list: the seam roots of every
    rank.  Both are tiny (de-duplicated on the device).  Returns (keys sorted unique, class_of_key, n_classes)."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    ks = [np.asarray(k, dtype=np.int64).ravel() for k in keys_list] + [np.asarray(edges, dtype=np.int64).ravel()]
    keys = np.unique(np.concatenate(ks)) if ks else np.zeros(0, np.int64)
    if keys.size == 0:
        return keys, np.zeros(0, np.int64), 0
    e = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
    a, b = np.searchsorted(keys, e[:, 0]), np.searchsorted(keys, e[:, 1])
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

    def to_torch(self, t):
        return t.contiguous()

    def from_torch(self, t):
        return t

    def seam_gid(self, L, valid, r0, off):
        """int32 [2,W]: slide-global index of the local root of the pixels of ext rows r0, r0+1 (-1 = not valid)"""
        g = L[r0:r0 + 2] + int(off)
        if valid is not None:
            g = self.torch.where(valid[r0:r0 + 2] != 0, g, self.torch.full_like(g, -1))
        return g.contiguous()

    def seam_export(self, L, valid, attr, off, He, has_top, has_bottom, nb_top):
        """One masked select + one device->host copy per round: rows (gid, neighbour gid or -1, attr) for the
        run starts of the rows shared with the neighbours.  nb_top: the lower rank's gid of my bottom rows."""
        t = self.torch
        parts = []
        for side, r0 in (("top", 0), ("bottom", He - 2)):
            if (side == "top" and not has_top) or (side == "bottom" and not has_bottom):
                continue
            g = L[r0:r0 + 2].to(t.int64) + int(off)
            if valid is not None:
                g = t.where(valid[r0:r0 + 2] != 0, g, t.full_like(g, -1))
            a = attr.view(-1)[(g - int(off)).clamp_(min=0)].to(t.int64) if attr is not None else t.zeros_like(g)
            nb = nb_top.to(t.int64) if (side == "bottom" and nb_top is not None) else t.full_like(g, -1)
            parts.append(t.stack([g.reshape(-1), nb.reshape(-1), a.reshape(-1)], dim=1))
        if not parts:
            return np.zeros((0, 3), np.int64)
        rows = t.cat(parts, dim=0)
        keep = rows[:, 0] >= 0
        keep[1:] &= (rows[1:, 0] != rows[:-1, 0]) | (rows[1:, 1] != rows[:-1, 1])
        return rows[keep].cpu().numpy()

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

    def relabel(self, L, keep, idmap, out=None):
        from ._cabi import check
        He, W = keep.shape
        labels = self.empty((He, W), "int32") if out is None else out
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


def _seam_round(be, comm, S, valid_of, attr_of, extra=None):
    """One reconciliation round.  For every local shard: gid of the rows shared with each neighbour (device),
    the lower rank ships its top rows up (point to point), the run starts of (own root, neighbour root, attr)
    are selected on the device and copied to the host in one piece; the tiny tables are all-gathered and
    every rank solves the same union on its host.  `extra(sh)` rides along in the same all-gather.
    Returns (keys, cls, ncls, entry keys, entry attrs, gathered extras)."""
    tops = [be.seam_gid(sh.L, valid_of(sh), 0, sh.off) if sh.has_top else None for sh in S]
    like = next((be.seam_gid(sh.L, valid_of(sh), sh.He - 2, sh.off) for sh in S if sh.has_bottom), None)
    from_lower = comm.send_up(be, tops, like)
    local = []
    for sh, nb in zip(S, from_lower):
        rows = be.seam_export(sh.L, valid_of(sh), attr_of(sh) if attr_of is not None else None, sh.off, sh.He,
                              sh.has_top, sh.has_bottom, nb)
        e = rows[rows[:, 1] >= 0][:, :2]
        assert e.size == 0 or e.min() >= 0, "seam pixels must be classified identically on both ranks"
        edges = np.unique(e, axis=0) if e.size else np.zeros((0, 2), np.int64)
        keys, first = np.unique(rows[:, 0], return_index=True)
        sh.seam_keys = keys
        local.append((edges, keys, rows[first, 2], extra(sh) if extra is not None else None))
    allv = comm.allgather(local)
    edges = np.concatenate([v[0] for v in allv]) if allv else np.zeros((0, 2), np.int64)
    keys, cls, ncls = seam_classes(edges, [v[1] for v in allv])
    ek = np.concatenate([v[1] for v in allv]) if allv else np.zeros(0, np.int64)
    ev = np.concatenate([v[2] for v in allv]) if allv else np.zeros(0, np.int64)
    return keys, cls, ncls, ek, ev, [v[3] for v in allv]


def postprocess_slide(shards, comm, H, W, be, direction_classes=9, min_area=20, radius=2, out_dtype="int64"):
    """Direction-aware post-processing (test_dam.py:455-563, postproc = 0) of an H x W slide whose rows
    are partitioned over comm.world ranks.

    shards: one dict per LOCAL rank (comm.local_ranks order) with the rank's OWN rows as numpy or
    backend arrays: dcm uint8 [T,Hl,W] (T = 1 or 8), prob float32 [3,Hl,W], point float32 [1,Hl,W].
    Returns the list of label arrays [Hl,W] (backend arrays) of the local ranks.  Raises the
    reference's AssertionError for a constant direction map."""
    G = comm.world
    parts = row_partition(H, G)
    import os
    import time
    _timing = os.environ.get("CDNET_SHARD_TIMING") is not None
    _marks = []

    def _mark(name):
        if _timing:
            import torch
            torch.cuda.synchronize()
            _marks.append((name, time.perf_counter()))

    _mark("start")
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
        if "dcm_ext" in d:
            # caller-allocated extended buffers (alloc_shard_buffers): own rows already sit at [lo, lo+Hl),
            # the ghost rows are filled in place -- no plane-sized copies anywhere
            sh.dcm_ext, sh.prob_ext, sh.point_ext = d["dcm_ext"], d["prob_ext"], d["point_ext"]
            assert sh.dcm_ext.shape[-2] == sh.He
        else:
            dcm, prob, point = asdev(d["dcm"]), asdev(d["prob"]), asdev(d["point"])
            assert dcm.shape[-2] == sh.Hl and dcm.shape[-1] == W
            sh.dcm_ext = be.empty((dcm.shape[0], sh.He, W), "uint8")
            sh.prob_ext = be.empty((3, sh.He, W), "float32")
            sh.point_ext = be.empty((1, sh.He, W), "float32")
            sh.dcm_ext[:, sh.lo:sh.lo + sh.Hl] = dcm
            sh.prob_ext[:, sh.lo:sh.lo + sh.Hl] = prob
            sh.point_ext[:, sh.lo:sh.lo + sh.Hl] = point
        sh.dcm = sh.dcm_ext[:, sh.lo:sh.lo + sh.Hl]
        sh.point = sh.point_ext[:, sh.lo:sh.lo + sh.Hl]
        S.append(sh)
    n_maps = int(S[0].dcm_ext.shape[0])

    def halo(get_rows, nrows):
        """first / last `nrows` own rows <-> the row neighbours, point to point, device resident"""
        up = [get_rows(sh, 0, nrows) if sh.has_top else None for sh in S]
        down = [get_rows(sh, sh.Hl - nrows, sh.Hl) if sh.has_bottom else None for sh in S]
        fu, fl = comm.exchange(be, up, down)
        return list(zip(fu, fl))

    def fill_ghosts(sh, ext, above, below):
        """write the received halo rows into the ghost rows of an extended plane (rows axis = -2)"""
        if above is not None:
            ext[..., 0:above.shape[-2], :] = above
        if below is not None:
            ext[..., ext.shape[-2] - below.shape[-2]:, :] = below

    # ---- phase 1: DDM codes (1-row class-map halo), global value flags and point maximum
    h_dcm = halo(lambda sh, a, b: sh.dcm[:, a:b], 1)
    h_pt = halo(lambda sh, a, b: sh.point[:, a:b], 1)
    loc = []
    for sh, (da, db), (pa, pb) in zip(S, h_dcm, h_pt):
        fill_ghosts(sh, sh.dcm_ext, da, db)
        fill_ghosts(sh, sh.point_ext, pa, pb)
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

    _mark('phase 2: boost')
    # ---- phase 2: boost + argmax on own rows, then 1-row halo of the inside mask
    for sh in S:
        # prob needs no halo (the boost is pointwise in prob); its ghost rows are never looked at
        sh.inside = be.boost(sh.codes, flags, sh.point_ext.reshape(sh.He, W), pmax, sh.prob_ext, n_maps)
        sh.inside_own = sh.inside[sh.lo:sh.lo + sh.Hl]
    h_in = halo(lambda sh, a, b: sh.inside_own[a:b], 1)
    for sh, (ia, ib) in zip(S, h_in):
        fill_ghosts(sh, sh.inside, ia, ib)

    def class_of(keys, cls, k):
        return cls[np.searchsorted(keys, k)]

    _mark('phase 3: forest')
    # ---- phase 3: forest of equal-value components; slide-global frame-touch flags for seam components
    for sh in S:
        sh.L, sh.touch = be.stage1(sh.inside, sh.rank == 0, sh.rank == G - 1)
    keys, cls, ncls, ek, ev, _ = _seam_round(be, comm, S, lambda sh: None, lambda sh: sh.touch)
    if ncls:
        ctouch = _class_reduce(keys, cls, ncls, ek, ev, "or")
        for sh in S:
            k = sh.seam_keys
            be.scatter(sh.touch, k - sh.off, ctouch[class_of(keys, cls, k)].astype(np.int32))

    _mark('phase 4: fill h')
    # ---- phase 4: fill holes, areas of own rows; slide-global areas for seam components
    for sh in S:
        sh.state, sh.area = be.stage2(sh.inside, sh.L, sh.touch, sh.lo, sh.lo + sh.Hl)
    keys, cls, ncls, ek, ev, _ = _seam_round(be, comm, S, lambda sh: sh.state, lambda sh: sh.area)
    if ncls:
        # every (rank, local root) contributes once; equal keys on two ranks are two local parts
        carea = np.minimum(_class_reduce(keys, cls, ncls, ek, ev, "sum"), 2 ** 31 - 1)
        for sh in S:
            k = sh.seam_keys
            be.scatter(sh.area, k - sh.off, carea[class_of(keys, cls, k)].astype(np.int32))

    _mark('phase 5: remove')
    # ---- phase 5: remove small, 8-connectivity; owners, excluded roots, numbering
    for sh in S:
        sh.keep = be.stage3(sh.state, sh.L, sh.area, min_area)
    keys, cls, ncls, ek, ev, _ = _seam_round(be, comm, S, lambda sh: sh.keep, None)
    croot = _class_reduce(keys, cls, ncls, keys, keys, "min") if ncls else np.zeros(0, np.int64)
    counts = []
    for sh in S:
        k = sh.seam_keys
        groot = croot[class_of(keys, cls, k)] if k.size else k
        own_lo, own_hi = sh.r0 * W, sh.r1 * W
        # a seam root is numbered by the rank that holds the class's first pixel in its OWN rows
        excl = (groot != k) | (k < own_lo) | (k >= own_hi)
        sh.seam_groot = groot
        sh.excluded = be.zeros((sh.He, W), "uint8")
        be.scatter(sh.excluded, (k - sh.off)[excl], np.ones(int(excl.sum()), np.uint8))
        sh.idmap, n_owned = be.stage4(sh.L, sh.keep, sh.excluded)
        counts.append(n_owned)
    tables = []
    for sh, n_owned in zip(S, counts):
        k, groot = sh.seam_keys, sh.seam_groot
        own = (groot == k) & (k >= sh.r0 * W) & (k < sh.r1 * W) if k.size else np.zeros(0, bool)
        ids = be.gather(sh.idmap, (k - sh.off)[own]) if own.any() else np.zeros(0, np.int64)
        tables.append((n_owned, k[own], np.asarray(ids).astype(np.int64)))
    g_tab = comm.allgather(tables)   # one exchange: per-rank owned-root counts + (seam class root, LOCAL id)
    offsets = np.concatenate([[0], np.cumsum([t[0] for t in g_tab])])
    tk = np.concatenate([t[1] for t in g_tab]) if g_tab else np.zeros(0, np.int64)
    tv = np.concatenate([t[2] + offsets[r] for r, t in enumerate(g_tab)]) if g_tab else np.zeros(0, np.int64)
    order = np.argsort(tk)
    tk, tv = tk[order], tv[order]
    for sh in S:
        be.add_scalar(sh.idmap, offsets[sh.rank])
        k, groot = sh.seam_keys, sh.seam_groot
        if k.size:
            pos = np.searchsorted(tk, groot)
            assert np.array_equal(tk[pos], groot), "every seam class must have exactly one owner"
            be.scatter(sh.idmap, k - sh.off, tv[pos].astype(np.int32))
        # labels of the extended tile land in the middle of a buffer with room for the extra halo rows of
        # the dilation (the ghost row already carries valid labels; rows beyond it come from the neighbour)
        r = int(radius)
        sh.pad_top = max(r - 1, 0) if sh.has_top else 0
        sh.pad_bot = max(r - 1, 0) if sh.has_bottom else 0
        sh.lab_big = be.empty((sh.pad_top + sh.He + sh.pad_bot, W), "int32")
        sh.labels = sh.lab_big[sh.pad_top:sh.pad_top + sh.He]
        be.relabel(sh.L, sh.keep, sh.idmap, out=sh.labels)

    _mark('phase 6: label ')
    # ---- phase 6: label dilation by disk(radius); rows beyond the ghost row come from the neighbour
    r = int(radius)
    outs = []
    if r > 1:
        own = lambda sh: sh.labels[sh.lo:sh.lo + sh.Hl]
        # the neighbour's own rows 2..r (seen from the seam): first send own[1:r] up / own[Hl-r:Hl-1] down
        up = [own(sh)[1:r] if sh.has_top else None for sh in S]
        down = [own(sh)[sh.Hl - r:sh.Hl - 1] if sh.has_bottom else None for sh in S]
        fu, fl = comm.exchange(be, up, down)
        for sh, a, b in zip(S, fu, fl):
            if a is not None:
                sh.lab_big[0:sh.pad_top] = a
            if b is not None:
                sh.lab_big[sh.pad_top + sh.He:] = b
    for sh in S:
        top = sh.pad_top + sh.lo
        out = be.dilate(sh.lab_big, r, out_dtype)[top:top + sh.Hl]
        outs.append(out)
    _mark("end")
    if _timing and S and S[0].rank == 0:
        print("shard timing (ms):", ", ".join("%s %.2f" % (_marks[i][0], 1e3 * (_marks[i + 1][1] - _marks[i][1]))
                                              for i in range(len(_marks) - 1)), flush=True)
    return outs


def alloc_shard_buffers(be, rank, world, H, W, n_maps):
    """Extended device buffers of one rank: dict(dcm_ext [T,He,W] u8, prob_ext [3,He,W] f32, point_ext
    [1,He,W] f32) plus views dcm / prob / point of the rank's OWN rows to be filled by the caller.  Passing
    the *_ext entries to postprocess_slide avoids every plane-sized copy."""
    r0, r1 = row_partition(H, world)[rank]
    lo = 1 if rank > 0 else 0
    He = (r1 - r0) + lo + (1 if rank < world - 1 else 0)
    d = {"dcm_ext": be.empty((n_maps, He, W), "uint8"), "prob_ext": be.empty((3, He, W), "float32"),
         "point_ext": be.empty((1, He, W), "float32")}
    for k in ("dcm", "prob", "point"):
        d[k] = d[k + "_ext"][:, lo:lo + (r1 - r0)]
    return d
