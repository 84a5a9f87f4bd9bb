"""Multi-GPU execution of the geometry hot path.

* Tile batches (BASELINE configs 2-4) are independent units: `shard_tiles` hands every rank its
  contiguous block, there is NO data-path collective (bench.py --gpus N).
* A whole-slide prediction map (config 5) is ROW-PARTITIONED: every rank post-processes its block of
  rows (test_dam.py:455-563 semantics for the whole slide, postproc = 0) and the ranks exchange only
    - halo rows (1 row of class map / point map / inside mask, radius-1 extra rows of labels), point to
      point between row neighbours,
    - two scalars (DDM value-present flags, point-map maximum),
    - per seam round a small TABLE of the run starts of the two rows shared with each neighbour
      (local root id, neighbour's root id, attribute): the tables are all-gathered and every rank solves
      the same union ON ITS DEVICE (csrc/seam.cu), which makes fill-holes (a hole is a background
      component that never reaches the SLIDE frame), remove-small (areas summed over shards) and the
      canonical raster-order numbering (ids = rank of the component's first pixel in the whole slide)
      identical to a single-GPU run.
  Nothing waits for the host between the first halo exchange and the final labels: every step is a
  kernel launch, an NCCL point-to-point / all-gather on device tensors, or a tiny device-side tensor op.
  The per-rank compute is the same sm_100a kernels cut into stages (cdnet_shard_* / cdnet_seam_* in
  include/cdnet_b200.h); the orchestration below is backend-agnostic so that it can be exercised on CPU
  with gloo (tests/sharded_numpy_backend.py is a numpy stand-in; the product has only the CUDA backend).

`comm` abstracts who is local: DistComm = one rank per process (torch.distributed: NCCL on the GPU box,
gloo in the CPU tests), SimComm = all ranks in this process (sharded == unsharded checks on one GPU).
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

    def allgather_tensor(self, be, tensors):
        """tensors[i]: same-shaped backend array of local rank i -> [G, ...] stack, one per local rank"""
        g = be.stack(tensors)
        return [g for _ in tensors]


class DistComm(object):
    """one rank per process over torch.distributed (NCCL with CUDA tensors, gloo on CPU)"""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.local_ranks = [self.rank]

    def _gather(self, t):
        """[...] per rank -> [G, ...]: ONE collective on a contiguous buffer (NCCL all-gather into a tensor; gloo and
        older back ends take the list form)"""
        out = t.new_empty((self.world,) + tuple(t.shape))
        try:
            self.dist.all_gather_into_tensor(out, t.contiguous(), group=self.group)
        except (RuntimeError, AttributeError, NotImplementedError):
            self.dist.all_gather([out[r] for r in range(self.world)], t.contiguous(), group=self.group)
        return out

    def exchange(self, be, up, down):
        """halo exchange with the two row neighbours as ONE all-gather: every rank contributes the rows for its upper
        and for its lower neighbour ([2, ...]; zeros where it has no neighbour) and picks its neighbours' slots.  The
        messages are a few rows (<= a few MB over all ranks), so the redundant data costs nothing on NVLink while one
        collective replaces a group of point-to-point sends and receives."""
        r, n = self.rank, self.world
        t_up = be.to_torch(up[0]) if up[0] is not None else None
        t_down = be.to_torch(down[0]) if down[0] is not None else None
        like = t_up if t_up is not None else t_down
        if like is None:
            return [None], [None]
        buf = like.new_zeros((2,) + tuple(like.shape))
        if t_up is not None:
            buf[0] = t_up
        if t_down is not None:
            buf[1] = t_down
        g = self._gather(buf)
        recv_up = g[r - 1, 1] if r > 0 else None      # the upper neighbour's bottom rows
        recv_down = g[r + 1, 0] if r < n - 1 else None  # the lower neighbour's top rows
        return ([be.from_torch(recv_up) if recv_up is not None else None],
                [be.from_torch(recv_down) if recv_down is not None else None])

    def send_up(self, be, tensors, like):
        """rank r > 0 hands tensors[0] to rank r-1; rank r < n-1 receives a tensor shaped `like` from r+1 (one all-gather)"""
        r, n = self.rank, self.world
        src = tensors[0] if tensors[0] is not None else like
        t = be.to_torch(src)
        if tensors[0] is None:
            t = t.new_zeros(t.shape)
        g = self._gather(t)
        return [be.from_torch(g[r + 1]) if r < n - 1 else None]

    def allgather_tensor(self, be, tensors):
        return [be.from_torch(self._gather(be.to_torch(tensors[0])))]


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

    def stack(self, tensors):
        return self.torch.stack([t for t in tensors], dim=0).contiguous()

    def to_torch(self, t):
        return t.contiguous()

    def from_torch(self, t):
        return t

    def _st(self):
        return self.torch.cuda.current_stream().cuda_stream

    # -- scalars of phase 1 stay on the device
    def pack_scalars(self, flags, pmax):
        return self.torch.cat([flags.reshape(1), pmax.reshape(1)]).contiguous()

    def combine_scalars(self, gathered):
        """[G,2] int32 (flags bits, order-preserving uint32 of the point max) -> (flags [1], pmax [1]) device"""
        t = self.torch
        flags = gathered[0, 0:1].clone()
        for r in range(1, gathered.shape[0]):
            flags |= gathered[r, 0:1]
        pm = (gathered[:, 1].to(t.int64) & 0xffffffff).max().reshape(1)
        pm = t.where(pm >= 2 ** 31, pm - 2 ** 32, pm).to(t.int32)
        return flags.contiguous(), pm.contiguous()

    def flags_to_host(self, flags):
        return int(flags.cpu().numpy().view(np.uint32)[0])

    def exclusive_offset(self, gathered_counts, rank):
        """[G,1] int32 per-rank counts -> [1] int32: sum over the lower ranks (device)"""
        c = gathered_counts.reshape(-1)
        return c[:rank].sum().to(self.torch.int32).reshape(1) if rank > 0 else self.zeros((1,), "int32")

    def add_offset(self, plane, offset):
        plane += offset

    # -- kernels
    def ddm_codes(self, dcm_ext, n_classes, row_lo, row_hi):
        from ._cabi import check
        T, He, W = dcm_ext.shape
        codes = self.empty((He, W), "uint16")
        flags = self.zeros((1,), "int32")
        check(self.L.cdnet_shard_ddm_codes(dcm_ext.data_ptr(), codes.data_ptr(), flags.data_ptr(), T, He, W,
                                           int(n_classes), int(row_lo), int(row_hi), self._st()), "shard_ddm_codes")
        return codes, flags

    def point_max(self, point_own):
        from ._cabi import check
        pm = self.zeros((1,), "int32")
        p = point_own.contiguous()
        check(self.L.cdnet_shard_point_max(p.data_ptr(), pm.data_ptr(), p.numel(), self._st()), "shard_point_max")
        return pm

    def boost(self, codes, flags, point_ext, pmax, prob_ext, n_maps):
        from ._cabi import check
        He, W = codes.shape
        inside = self.empty((He, W), "uint8")
        status = self.zeros((1,), "int32")
        check(self.L.cdnet_shard_boost(codes.data_ptr(), flags.data_ptr(), point_ext.data_ptr(), pmax.data_ptr(),
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
        """-> (idmap, n_owned [1] int32 on the device)"""
        from ._cabi import check
        He, W = keep.shape
        idmap = self.zeros((He, W), "int32")  # only root pixels are written; add_offset reads the whole plane
        rowcnt = self.empty((He,), "int32")
        n = self.zeros((1,), "int32")
        check(self.L.cdnet_shard_label_stage4(L.data_ptr(), keep.data_ptr(), excluded.data_ptr(), idmap.data_ptr(),
                                              rowcnt.data_ptr(), n.data_ptr(), He, W, self._st()), "stage4")
        return idmap, n

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

    # -- postproc = 1: process() on own rows + overlap, slide-global marker ids
    def ws_process(self, inside_ext, own_lo, own_hi, min_size):
        """-> (labels int32 [He,W] with tile-local marker ids, rowmax int32 [He], status int32 [1])"""
        from ._cabi import check
        He, W = inside_ext.shape
        labels = self.empty((He, W), "int32")
        rowmax = self.empty((He,), "int32")
        status = self.empty((1,), "int32")
        ws = self.api._workspace(self.L.cdnet_ws_postproc_workspace_bytes(1, He, W), self.dev)
        check(self.L.cdnet_shard_ws_process(inside_ext.data_ptr(), labels.data_ptr(), rowmax.data_ptr(),
                                            status.data_ptr(), He, W, int(own_lo), int(own_hi), int(min_size),
                                            ws.data_ptr(), ws.numel(), self._st()), "shard_ws_process")
        return labels, rowmax, status

    def marker_ranges(self, rowmax, own_lo, own_hi):
        """markers that start above the own rows, markers that start on them: two [1] int32 device tensors"""
        t = self.torch
        a = rowmax[:own_lo].max().reshape(1) if own_lo > 0 else self.zeros((1,), "int32")
        b = t.maximum(a, rowmax[:own_hi].max().reshape(1))
        return a, b - a

    def ws_own_ids(self, labels, scalars):
        """slide-global id of the markers this rank owns (local ids above+1 .. above+owned), 0 elsewhere; scalars =
        [above, owned, offset] on the device"""
        t = self.torch
        above, owned, off = scalars[0], scalars[1], scalars[2]
        mine = (labels > above) & (labels <= above + owned)
        return t.where(mine, labels - above + off, t.zeros_like(labels))

    def ws_adopt(self, lut, local_rows, global_rows):
        """lut[local id] = the owner's id, read where the neighbour's own rows overlap this tile"""
        t = self.torch
        l, g = local_rows.reshape(-1), global_rows.reshape(-1)
        ok = (l > 0) & (g > 0)
        zero = t.zeros_like(l)
        lut.index_copy_(0, t.where(ok, l, zero).long(), t.where(ok, g, zero))   # entry 0 stays 0

    def ws_relabel(self, labels_own, scalars, lut, out):
        """final labels of the own rows into `out` (a [Hl,W] view); -> [1] int32, 1 if a label found no owner"""
        from ._cabi import check
        Hl, W = labels_own.shape
        assert labels_own.is_contiguous() and out.is_contiguous()
        err = self.zeros((1,), "int32")
        check(self.L.cdnet_shard_ws_relabel(labels_own.data_ptr(), scalars.data_ptr(), lut.data_ptr(), out.data_ptr(),
                                            err.data_ptr(), Hl, W, self._st()), "shard_ws_relabel")
        return err

    def pack_status_ws(self, flags, ws_status, unresolved):
        """[flags word, error bits] like pack_status: 4 = a nucleus reaches beyond the overlap, 8 = no background,
        16 = unresolved marker id, 32 = watershed queue overflow"""
        t = self.torch
        from ._cabi import S_NO_BACKGROUND, S_SHARD_OVERFLOW, S_WS_OVERFLOW
        w = ws_status.reshape(-1)[:1]
        err = (((w & S_SHARD_OVERFLOW) != 0).to(t.int32) * 4 + ((w & S_NO_BACKGROUND) != 0).to(t.int32) * 8 +
               unresolved.reshape(-1)[:1] * 16 + ((w & S_WS_OVERFLOW) != 0).to(t.int32) * 32)
        return t.cat([flags.reshape(-1)[:1].to(t.int32), err.to(t.int32)])

    # -- seam rounds (csrc/seam.cu)
    def seam_init(self, sh, world, W):
        sh.cap = 4 * W + 1
        sh.emitted = self.zeros((sh.He, W), "int32")
        nb = self.L.cdnet_seam_workspace_bytes(world, sh.cap)
        assert nb > 0, "slide too wide for the seam tables"
        sh.seam_ws = self.empty((nb,), "uint8")
        sh.world = world

    def seam_gid(self, L, valid, r0, off):
        """int32 [2,W]: slide-global index of the local root of the pixels of ext rows r0, r0+1 (-1 = not valid)"""
        g = L[r0:r0 + 2] + int(off)
        if valid is not None:
            g = self.torch.where(valid[r0:r0 + 2] != 0, g, self.torch.full_like(g, -1))
        return g.contiguous()

    def seam_export(self, sh, valid, attr, round_id, nb_gid):
        from ._cabi import check
        tbl = self.zeros((sh.cap, 4), "int32")  # the whole table travels (all-gather): no undefined tail
        check(self.L.cdnet_seam_export(sh.L.data_ptr(), valid.data_ptr() if valid is not None else None,
                                       attr.data_ptr() if attr is not None else None, sh.emitted.data_ptr(),
                                       int(round_id), int(sh.off), sh.He, sh.L.shape[1], 1 if sh.has_top else 0,
                                       1 if sh.has_bottom else 0, nb_gid.data_ptr() if nb_gid is not None else None,
                                       tbl.data_ptr(), sh.cap, self._st()), "seam_export")
        return tbl

    def seam_solve(self, sh, gathered, mode, W, plane=None, excluded=None):
        from ._cabi import check
        check(self.L.cdnet_seam_solve(gathered.data_ptr(), sh.world, sh.cap, sh.rank, int(mode), int(sh.off),
                                      sh.r0 * W, sh.r1 * W, plane.data_ptr() if plane is not None else None,
                                      excluded.data_ptr() if excluded is not None else None, sh.seam_ws.data_ptr(),
                                      sh.seam_ws.numel(), self._st()), "seam_solve")

    def seam_ids_export(self, sh, gathered, idmap, round_id, W):
        from ._cabi import check
        tbl2 = self.zeros((sh.cap, 4), "int32")
        check(self.L.cdnet_seam_ids_export(gathered.data_ptr(), sh.world, sh.cap, sh.rank, int(sh.off), sh.r0 * W,
                                           sh.r1 * W, idmap.data_ptr(), sh.emitted.data_ptr(), int(round_id),
                                           tbl2.data_ptr(), sh.seam_ws.data_ptr(), sh.seam_ws.numel(), self._st()),
              "seam_ids_export")
        return tbl2

    def seam_ids_apply(self, sh, gathered, gathered2, idmap):
        from ._cabi import check
        check(self.L.cdnet_seam_ids_apply(gathered.data_ptr(), gathered2.data_ptr(), sh.world, sh.cap, sh.rank,
                                          int(sh.off), idmap.data_ptr(), sh.seam_ws.data_ptr(), sh.seam_ws.numel(),
                                          self._st()), "seam_ids_apply")

    def seam_errors(self, tables):
        """error bits of the exported tables (one device->host read, at the very end)"""
        return int(self.torch.stack([t[0, 1] for t in tables]).max().item()) if tables else 0

    def pack_status(self, flags, tables):
        """[flags word, max error word of the seam tables] as ONE device tensor (enqueue only)"""
        t = self.torch
        err = t.stack([x[0, 1] for x in tables]).max().reshape(1).to(t.int32) if tables else self.zeros((1,), "int32")
        return t.cat([flags.reshape(-1)[:1].to(t.int32), err])

    def status_to_host(self, packed):
        """the step's single device -> host read: (flags as uint32, seam error bits)"""
        h = packed.cpu().numpy()
        return int(h[:1].view(np.uint32)[0]), int(h[1])


# --------------------------------------------------------------------------------------------------
# whole-slide post-processing
# --------------------------------------------------------------------------------------------------
class _Shard(object):
    pass


def _cat_bytes(arrays):
    """[array, ...] (torch tensors or numpy arrays of any dtype) -> one flat uint8 array with the bytes of all"""
    if isinstance(arrays[0], np.ndarray):
        return np.concatenate([np.ascontiguousarray(a).view(np.uint8).reshape(-1) for a in arrays])
    import torch
    return torch.cat([a.contiguous().view(torch.uint8).reshape(-1) for a in arrays])


def _split_bytes(buf, like):
    """inverse of _cat_bytes: views of `buf` shaped and typed like the arrays of `like`"""
    out, off = [], 0
    for a in like:
        if isinstance(a, np.ndarray):
            n = a.size * a.dtype.itemsize
            out.append(buf[off:off + n].view(a.dtype).reshape(a.shape))
        else:
            n = a.numel() * a.element_size()
            out.append(buf[off:off + n].view(a.dtype).reshape(a.shape))
        off += n
    return out


def _seam_tables(be, comm, S, valid_of, attr_of, round_id):
    """export + all-gather of one seam round; returns the gathered table of every local shard"""
    tops = [be.seam_gid(sh.L, valid_of(sh), 0, sh.off) if sh.has_top else None for sh in S]
    like = next((be.seam_gid(sh.L, valid_of(sh), sh.He - 2, sh.off) for sh in S if sh.has_bottom), None)
    from_lower = comm.send_up(be, tops, like)
    tables = [be.seam_export(sh, valid_of(sh), attr_of(sh) if attr_of is not None else None, round_id, nb)
              for sh, nb in zip(S, from_lower)]
    for sh, t in zip(S, tables):
        sh.tables.append(t)
    return comm.allgather_tensor(be, tables)


def _check_flags(flags, n_maps):
    for t in range(n_maps):
        f = (flags >> (3 * t)) & 7
        if f in (0, 1, 2, 4):
            raise AssertionError("constant direction map: generate_dd_map is NaN (test_dam.py:535)")


def _watershed_tail(S, comm, be, W, r, out_dtype, K, halo, _mark, n_maps):
    """postproc = 1 after the boost: inside mask of the own rows -> process() -> dilation (test_dam.py:559-563).

    Why overlap + ownership is exact: process() treats every 4-connected component of the mask on its own (distance
    normalised by the component's maximum :24-26, markers inside it :39-46, flooding confined to the mask :47), apart
    from the marker numbering of :44, which is the raster order of the markers' first pixels over the slide.  A
    component that touches a rank's rows and stays inside its extended tile is therefore segmented there exactly as on
    the whole slide, with tile-local ids; the kernel reports when a component does not stay inside.  Ids: a marker
    belongs to the rank whose own rows hold its first pixel; tile-local ids are raster ordered too, so the owner's
    slide-global id is (markers owned by lower ranks) + (rank among its own).  Labels of the own rows that belong to a
    neighbour's marker are read off the neighbour's labels on the overlap rows, where both tiles show the same
    regions."""
    _mark("phase 3: watershed")
    h_in = halo(lambda sh, a, b: sh.inside_own[a:b], K)
    counts = []
    for sh, (ia, ib) in zip(S, h_in):
        sh.kt = ia.shape[-2] if ia is not None else 0
        kb = ib.shape[-2] if ib is not None else 0
        sh.inw = be.empty((sh.kt + sh.Hl + kb, W), "uint8")
        sh.inw[sh.kt:sh.kt + sh.Hl] = sh.inside_own
        if ia is not None:
            sh.inw[:sh.kt] = ia
        if ib is not None:
            sh.inw[sh.kt + sh.Hl:] = ib
        # min_size 10: what the DAM path passes to process() (test_dam.py:559)
        sh.wl, rowmax, sh.wstat = be.ws_process(sh.inw, sh.kt, sh.kt + sh.Hl, 10)
        sh.n_above, sh.n_owned = be.marker_ranges(rowmax, sh.kt, sh.kt + sh.Hl)
        counts.append(sh.n_owned)
    g_cnt = comm.allgather_tensor(be, counts)
    _mark("phase 4: marker ids")
    top = lambda sh, a, b: slice(sh.kt + a, sh.kt + b)
    for sh, gc in zip(S, g_cnt):
        sh.scal = be.stack([sh.n_above.reshape(()), sh.n_owned.reshape(()), be.exclusive_offset(gc, sh.rank).reshape(())])
    # the owners' ids on the K rows either side of every seam (only those rows are ever looked at)
    h_ids = halo(lambda sh, a, b: be.ws_own_ids(sh.wl[top(sh, a, b)], sh.scal), K)
    packed = []
    for sh, (la, lb) in zip(S, h_ids):
        lut = be.zeros((sh.wl.shape[0] * W // 2 + 2,), "int32")   # more entries than 4-connected components
        if la is not None:
            be.ws_adopt(lut, sh.wl[:sh.kt], la)
        if lb is not None:
            be.ws_adopt(lut, sh.wl[sh.kt + sh.Hl:], lb)
        # final labels land in the middle of the buffer the dilation reads, with room for the neighbours' rows
        sh.pt = r if sh.has_top else 0
        sh.big = be.empty((sh.pt + sh.Hl + (r if sh.has_bottom else 0), W), "int32")
        sh.final = sh.big[sh.pt:sh.pt + sh.Hl]
        unresolved = be.ws_relabel(sh.wl[sh.kt:sh.kt + sh.Hl], sh.scal, lut, sh.final)
        packed.append(be.pack_status_ws(sh.flags, sh.wstat, unresolved))
        del lut
    g_stat = comm.allgather_tensor(be, packed)
    _mark("phase 6: dilation")
    outs = []
    h_lab = halo(lambda sh, a, b: sh.final[a:b], r) if r > 0 else [(None, None)] * len(S)
    for sh, (la, lb) in zip(S, h_lab):
        if la is not None:
            sh.big[:sh.pt] = la
        if lb is not None:
            sh.big[sh.pt + sh.Hl:] = lb
        outs.append(be.dilate(sh.big, r, out_dtype)[sh.pt:sh.pt + sh.Hl])
    _mark("end")

    def check():
        # every rank looks at the error bits of ALL ranks, so that either all of them raise or none does
        h = np.asarray(be.to_host(g_stat[0])).reshape(-1, 2)
        _check_flags(int(h[:1, 0].astype(np.int32).view(np.uint32)[0]), n_maps)
        err = int(np.bitwise_or.reduce(h[:, 1]))
        if err & 8:
            raise ValueError("the mask of a shard has no background pixel (postproc_other.py:18-19 raises there)")
        if err & 4:
            raise RuntimeError("whole-slide watershed: a nucleus reaches beyond the %d overlap rows of a shard; "
                               "raise `overlap` (or use fewer ranks)" % K)
        if err & 16:
            raise RuntimeError("whole-slide watershed: a label of a shard found no owner")
        if err & 32:
            raise RuntimeError("watershed queue overflow")
    return outs, check


def postprocess_slide(shards, comm, H, W, be, direction_classes=9, min_area=20, radius=2, out_dtype=None,
                      timings=None, defer_checks=False, postproc=0, overlap=128):
    """Direction-aware post-processing (test_dam.py:455-563) of an H x W slide whose rows are partitioned over
    comm.world ranks.  postproc = 0: fill holes / remove small / label (exact for any component, seam union-find).
    postproc = 1: process() = marker-controlled watershed (test_dam.py:559, postproc_other.py:32-48); every rank
    runs it on its own rows plus `overlap` rows of each neighbour and the ranks agree on slide-global marker ids --
    exact as long as no nucleus that touches a rank's rows reaches beyond its overlap rows, RuntimeError otherwise
    (the distance normalisation of a nucleus needs the whole nucleus).  out_dtype defaults to the reference's: int64
    for postproc 0, int32 for postproc 1.

    shards: one dict per LOCAL rank (comm.local_ranks order), either with the rank's OWN rows as numpy or
    backend arrays -- dcm uint8 [T,Hl,W] (T = 1 or 8), prob float32 [3,Hl,W], point float32 [1,Hl,W] -- or
    with the extended buffers of `alloc_shard_buffers` (no plane-sized copies).
    Returns the list of label arrays [Hl,W] (backend arrays) of the local ranks.  Raises the
    reference's AssertionError for a constant direction map (checked once, at the end).
    timings: optional dict; when given, the device is synchronised after every phase and the dict receives
    {phase name: milliseconds} (a diagnostic mode: the synchronisations cost time themselves).
    defer_checks: return (labels, check) instead, where check() performs the one host round trip (status of the whole
    slide) and raises like the plain call -- everything before it is enqueue-only, so it can be captured in a CUDA
    graph (SlidePlan)."""
    G = comm.world
    parts = row_partition(H, G)
    postproc = int(postproc)
    if postproc not in (0, 1):
        raise ValueError("whole-slide post-processing supports postproc 0 and 1")
    if out_dtype is None:
        out_dtype = "int64" if postproc == 0 else "int32"
    if postproc == 1 and G > 1 and min(b - a for a, b in parts) < max(int(overlap), int(radius), 1):
        raise ValueError("whole-slide watershed: shards need at least overlap = %d rows each (H = %d over %d ranks)"
                         % (int(overlap), H, G))
    if G > 1 and min(b - a for a, b in parts) < max(2, int(radius)):
        # phase 6 fetches radius-1 label rows from each row neighbour only: a shard thinner than the radius would need
        # rows from two shards away
        raise ValueError("whole-slide shards need at least max(2, radius) = %d rows each (H = %d over %d ranks)"
                         % (max(2, int(radius)), H, G))
    import os
    import time
    _timing = timings is not None or os.environ.get("CDNET_SHARD_TIMING") is not None
    _marks = []

    def _mark(name):
        if _timing:
            import torch
            torch.cuda.synchronize()
            _marks.append((name, time.perf_counter()))

    def _report():
        if timings is not None:
            for i in range(len(_marks) - 1):
                timings[_marks[i][0]] = 1e3 * (_marks[i + 1][1] - _marks[i][1])
        if _timing and timings is None and S and S[0].rank == 0:
            print("shard timing (ms):", ", ".join("%s %.2f" % (_marks[i][0], 1e3 * (_marks[i + 1][1] - _marks[i][1]))
                                                  for i in range(len(_marks) - 1)), flush=True)

    _mark("phase 0: set-up")
    S = []
    for rank, d in zip(comm.local_ranks, shards):
        sh = _Shard()
        sh.rank, (sh.r0, sh.r1) = rank, parts[rank]
        sh.has_top, sh.has_bottom = rank > 0, rank < G - 1
        sh.Hl = sh.r1 - sh.r0
        sh.lo = 1 if sh.has_top else 0            # ext row of the first own row
        sh.He = sh.Hl + sh.lo + (1 if sh.has_bottom else 0)
        sh.off = (sh.r0 - sh.lo) * W              # slide-global index of ext pixel 0
        sh.tables = []
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
            sh.prob_ext = be.zeros((3, sh.He, W), "float32")
            sh.point_ext = be.empty((1, sh.He, W), "float32")
            sh.dcm_ext[:, sh.lo:sh.lo + sh.Hl] = dcm
            sh.prob_ext[:, sh.lo:sh.lo + sh.Hl] = prob
            sh.point_ext[:, sh.lo:sh.lo + sh.Hl] = point
        sh.dcm = sh.dcm_ext[:, sh.lo:sh.lo + sh.Hl]
        sh.point = sh.point_ext[:, sh.lo:sh.lo + sh.Hl]
        be.seam_init(sh, G, W)
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

    # ---- phase 1: DDM codes (1-row class-map halo), slide-global value flags and point maximum
    # the class-map row and the point-map row travel in ONE message (bytes of both, concatenated)
    def both_rows(sh, a, b):
        return _cat_bytes([sh.point[:, a:b], sh.dcm[:, a:b]])  # the float rows first: their view stays 4-byte aligned
    h_both = halo(both_rows, 1)
    h_dcm, h_pt = [], []
    for sh, (above, below) in zip(S, h_both):
        like = [sh.point[:, 0:1], sh.dcm[:, 0:1]]
        ua = _split_bytes(above, like) if above is not None else (None, None)
        ub = _split_bytes(below, like) if below is not None else (None, None)
        h_pt.append((ua[0], ub[0]))
        h_dcm.append((ua[1], ub[1]))
    loc = []
    for sh, (da, db), (pa, pb) in zip(S, h_dcm, h_pt):
        fill_ghosts(sh, sh.dcm_ext, da, db)
        fill_ghosts(sh, sh.point_ext, pa, pb)
        sh.codes, fl = be.ddm_codes(sh.dcm_ext, direction_classes, sh.lo, sh.lo + sh.Hl)
        loc.append(be.pack_scalars(fl, be.point_max(sh.point)))
    g_scal = comm.allgather_tensor(be, loc)

    _mark("phase 2: boost")
    # ---- phase 2: boost + argmax on own rows, then 1-row halo of the inside mask
    for sh, gs in zip(S, g_scal):
        sh.flags, pmax = be.combine_scalars(gs)
        # prob needs no halo (the boost is pointwise in prob); its ghost rows are never looked at
        sh.inside = be.boost(sh.codes, sh.flags, sh.point_ext.reshape(sh.He, W), pmax, sh.prob_ext, n_maps)
        sh.inside_own = sh.inside[sh.lo:sh.lo + sh.Hl]
    if postproc == 1:
        outs, check = _watershed_tail(S, comm, be, W, int(radius), out_dtype, int(overlap), halo, _mark, n_maps)
        _report()
        if defer_checks:
            return outs, check
        check()
        return outs
    h_in = halo(lambda sh, a, b: sh.inside_own[a:b], 1)
    for sh, (ia, ib) in zip(S, h_in):
        fill_ghosts(sh, sh.inside, ia, ib)

    _mark("phase 3: forest")
    # ---- phase 3: forest of equal-value components; slide-global frame-touch flags for seam components
    for sh in S:
        sh.L, sh.touch = be.stage1(sh.inside, sh.rank == 0, sh.rank == G - 1)
    g_tab = _seam_tables(be, comm, S, lambda sh: None, lambda sh: sh.touch, 1)
    for sh, g in zip(S, g_tab):
        be.seam_solve(sh, g, 0, W, plane=sh.touch)

    _mark("phase 4: fill holes")
    # ---- phase 4: fill holes, areas of own rows; slide-global areas for seam components
    for sh in S:
        sh.state, sh.area = be.stage2(sh.inside, sh.L, sh.touch, sh.lo, sh.lo + sh.Hl)
    g_tab = _seam_tables(be, comm, S, lambda sh: sh.state, lambda sh: sh.area, 2)
    for sh, g in zip(S, g_tab):
        be.seam_solve(sh, g, 1, W, plane=sh.area)

    _mark("phase 5: remove small")
    # ---- phase 5: remove small, 8-connectivity; owners, excluded roots, numbering
    for sh in S:
        sh.keep = be.stage3(sh.state, sh.L, sh.area, min_area)
    g_tab = _seam_tables(be, comm, S, lambda sh: sh.keep, None, 3)
    counts = []
    for sh, g in zip(S, g_tab):
        # a seam class is numbered by the rank that holds its first pixel (smallest gid) in its OWN rows
        sh.excluded = be.zeros((sh.He, W), "uint8")
        be.seam_solve(sh, g, 2, W, excluded=sh.excluded)
        sh.idmap, n_owned = be.stage4(sh.L, sh.keep, sh.excluded)
        counts.append(n_owned)
    g_cnt = comm.allgather_tensor(be, counts)
    tabs2 = []
    for sh, g, gc in zip(S, g_tab, g_cnt):
        be.add_offset(sh.idmap, be.exclusive_offset(gc, sh.rank))   # ids = ranks below + local raster rank
        tabs2.append(be.seam_ids_export(sh, g, sh.idmap, 4, W))      # owners publish (class root, final id)
    g_tab2 = comm.allgather_tensor(be, tabs2)
    r = int(radius)
    for sh, g, g2 in zip(S, g_tab, g_tab2):
        be.seam_ids_apply(sh, g, g2, sh.idmap)
        # labels of the extended tile land in the middle of a buffer with room for the extra halo rows of
        # the dilation (the ghost row already carries valid labels; rows beyond it come from the neighbour)
        sh.pad_top = max(r - 1, 0) if sh.has_top else 0
        sh.pad_bot = max(r - 1, 0) if sh.has_bottom else 0
        sh.lab_big = be.empty((sh.pad_top + sh.He + sh.pad_bot, W), "int32")
        sh.labels = sh.lab_big[sh.pad_top:sh.pad_top + sh.He]
        be.relabel(sh.L, sh.keep, sh.idmap, out=sh.labels)

    _mark("phase 6: dilation")
    # ---- phase 6: label dilation by disk(radius); rows beyond the ghost row come from the neighbour
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
    _report()
    # ---- the only host round trip: status of the whole slide (one small tensor, packed while enqueueing)
    packed = [be.pack_status(sh.flags, sh.tables) for sh in S] if hasattr(be, "pack_status") else None

    def check():
        if packed is not None:
            st = [be.status_to_host(p) for p in packed]
            flags, err = st[0][0], max(e for _, e in st)
        else:
            flags = be.flags_to_host(S[0].flags)
            err = max(be.seam_errors(sh.tables) for sh in S)
        _check_flags(flags, n_maps)
        if err & 1:
            raise RuntimeError("seam pixels were classified differently by two neighbouring ranks")
        if err & 2:
            raise RuntimeError("seam table overflow")
    if defer_checks:
        return outs, check
    check()
    return outs


class SlidePlan(object):
    """One rank's whole-slide post-processing step captured in a CUDA graph.

    A step is ~100 kernel launches, a dozen small torch ops and ~10 NCCL all-gathers; issued from Python it costs
    about a millisecond of host time per step and leaves gaps on the device between the short kernels of a shard.
    The shapes are fixed for a given slide, so the whole enqueue sequence (collectives included: NCCL calls are
    capturable) is recorded once and replayed.  `bufs` are the rank's extended buffers (alloc_shard_buffers); fill the
    dcm / prob / point views, call run(), read `labels` ([Hl, W], the rank's own rows).  CUDA backend only."""

    def __init__(self, bufs, comm, H, W, be, direction_classes=9, min_area=20, radius=2, out_dtype=None, postproc=0,
                 overlap=128):
        import torch
        self.torch, self.be = torch, be
        args = ([bufs], comm, H, W, be, direction_classes, min_area, radius, out_dtype)
        kw = dict(postproc=postproc, overlap=overlap)
        postprocess_slide(*args, **kw)  # eager warm-up: scratch memory, kernel attributes, NCCL channels
        torch.cuda.synchronize()
        torch.cuda.empty_cache()  # the graph's private pool holds the step's intermediates from here on
        self.graph = torch.cuda.CUDAGraph()
        # thread_local: the NCCL watchdog thread may query events while this thread captures
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            outs, self._check = postprocess_slide(*args, defer_checks=True, **kw)
        self.labels = outs[0]

    def launch(self):
        self.graph.replay()

    def run(self):
        self.graph.replay()
        self._check()
        return self.labels


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
    # the ghost rows of prob are never exchanged (the boost is pointwise in prob and its result on ghost rows is replaced by
    # the neighbour's): give them a defined value once, so that no kernel ever reads uninitialised memory
    d["prob_ext"][:, :lo] = 0
    d["prob_ext"][:, lo + (r1 - r0):] = 0
    return d
