"""Instance metrics of the reference (stats_utils.py) on top of one CUDA reduction.

Drop-ins with the reference's names, arguments, return structure and error behaviour:

    get_fast_aji(true, pred)            stats_utils.py:7-98     (called at test.py:342, test_dam.py:616)
    get_fast_aji_plus(true, pred)       stats_utils.py:101-177
    get_fast_pq(true, pred, match_iou)  stats_utils.py:182-275  (test.py:345, test_dam.py:621)
    get_fast_dice_2(true, pred)         stats_utils.py:279-318
    get_dice_1(true, pred)              stats_utils.py:324-334  (test.py:343, test_dam.py:617)
    get_dice_2(true, pred)              stats_utils.py:338-357
    remap_label(pred, by_size=False)    stats_utils.py:361-389

The reference builds one H x W mask per instance and loops over instance pairs.  Here `cdnet_label_pairs`
(csrc/metrics.cu) reduces the two label images to the sparse table n[t][q] of pixel counts in one pass on the
GPU; what is left for the host is float64 arithmetic on a few thousand pairs, written in the reference's order of
operations so that the scores are the same float64 values.  There is no CPU path for the reduction.
"""
import numpy as np
import torch

from . import _cabi
from ._cabi import check, CdnetError
from .api import _device, _workspace, _stream, _ptr

S_PAIR_OVERFLOW, S_PAIR_RANGE = 4, 8


class PairTable(object):
    """Distinct (true id, pred id) pairs of one tile with their pixel counts, sorted by (true, pred)."""

    def __init__(self, keys, counts, n_pixels=None):
        keys = np.asarray(keys, dtype=np.uint64)
        order = np.argsort(keys, kind="stable")
        keys = keys[order]
        self.t = (keys >> np.uint64(32)).astype(np.int64)
        self.q = (keys & np.uint64(0xffffffff)).astype(np.int64)
        self.n = np.asarray(counts, dtype=np.int64)[order]
        if n_pixels is not None and int(self.n.sum()) != int(n_pixels):
            raise CdnetError("pair table does not add up to the tile size (%d != %d)" % (int(self.n.sum()), n_pixels))
        mt = int(self.t.max()) if self.t.size else 0
        mq = int(self.q.max()) if self.q.size else 0
        self.area_t = np.bincount(self.t, weights=self.n, minlength=mt + 1).astype(np.int64)
        self.area_q = np.bincount(self.q, weights=self.n, minlength=mq + 1).astype(np.int64)
        self.true_ids = np.flatnonzero(self.area_t)     # ascending == np.unique order
        self.pred_ids = np.flatnonzero(self.area_q)
        both = (self.t > 0) & (self.q > 0)
        self.ot, self.oq, self.on = self.t[both], self.q[both], self.n[both]   # the overlapping instance pairs

    def instance_lists(self):
        """The reference's `true_id_list[1:]` / `pred_id_list[1:]` (stats_utils.py:17-28): it drops the smallest
        id as "the background" and indexes its mask lists by id, i.e. needs ids 0, 1..N without gaps."""
        def one(ids, what):
            inst = ids[1:] if ids.size else ids
            if ids.size and (int(ids[0]) != 0 or (inst.size and int(inst[-1]) != inst.size)):
                raise IndexError("%s: instance ids must be contiguous 1..N with background 0 present "
                                 "(stats_utils.py:10-12: call remap_label first)" % what)
            return inst
        return one(self.true_ids, "true"), one(self.pred_ids, "pred")


def label_pairs_cuda(true, pred, cap=None):
    """true, pred: int32 / int64 tensors [B,H,W] on the GPU -> list of PairTable (one per tile)."""
    L = _cabi.lib()
    dev = _device(true.device)
    if true.shape != pred.shape or true.dim() != 3:
        raise ValueError("true and pred must both be [B,H,W]")
    if true.dtype != pred.dtype or true.dtype not in (torch.int32, torch.int64):
        true, pred = true.to(torch.int64), pred.to(torch.int64)
    true, pred = true.contiguous(), pred.contiguous()
    B, H, W = true.shape
    eb = 4 if true.dtype == torch.int32 else 8
    full = H * W + 1                                     # distinct pairs can never exceed the pixel count
    cap = int(cap) if cap else min(full, max(4096, (H * W) // 32))
    while True:
        # zero-filled: tiles with fewer pairs than the fullest one leave their tail entries unwritten, and the
        # rectangular device -> host copy below would otherwise read uninitialised memory (compute-sanitizer initcheck)
        keys = torch.zeros((B, cap), dtype=torch.int64, device=dev)
        counts = torch.zeros((B, cap), dtype=torch.int32, device=dev)
        n_out = torch.empty((B,), dtype=torch.int32, device=dev)
        status = torch.empty((B,), dtype=torch.int32, device=dev)
        ws = _workspace(L.cdnet_label_pairs_workspace_bytes(B, cap), dev)
        check(L.cdnet_label_pairs(_ptr(true), _ptr(pred), eb, _ptr(keys), _ptr(counts), _ptr(n_out), _ptr(status),
                                  B, H, W, cap, _ptr(ws), ws.numel(), _stream()), "cdnet_label_pairs")
        st = status.cpu().numpy()
        if (st & S_PAIR_RANGE).any():
            raise ValueError("label ids must lie in [0, 2^31)")
        if not (st & S_PAIR_OVERFLOW).any():
            break
        if cap >= full:
            raise CdnetError("pair table overflow with cap == H*W + 1 (cannot happen)")
        cap = full                                       # rare: retry with the guaranteed bound
    n = n_out.cpu().numpy()
    m = int(n.max()) if B else 0
    hk = keys[:, :m].cpu().numpy().view(np.uint64)
    hc = counts[:, :m].cpu().numpy()
    return [PairTable(hk[b, :n[b]], hc[b, :n[b]], H * W) for b in range(B)]


def _to_device(a):
    a = np.asarray(a)
    if a.ndim != 2:
        raise ValueError("expected a 2-D label image")
    if a.dtype not in (np.int32, np.int64):
        if a.dtype.kind not in "iub":
            raise TypeError("label images must have an integer dtype")
        a = a.astype(np.int64)
    return torch.from_numpy(np.ascontiguousarray(a)).to(_device(None))[None]


def pair_table(true, pred):
    """numpy [H,W] label images -> PairTable."""
    true, pred = np.asarray(true), np.asarray(pred)
    if true.shape != pred.shape:
        raise ValueError("true and pred differ in shape")
    t, p = _to_device(true), _to_device(pred)
    return label_pairs_cuda(t, p)[0]


# ---- float64 epilogues on the pair table ---------------------------------------------------------------------
def _best_pred_per_true(T, iou):
    """first maximum of each row of the reference's dense iou matrix (np.argmax, stats_utils.py:63): among the
    overlapping preds of a true id the largest iou, ties -> smallest pred id.  Returns indices into T.o*."""
    if T.ot.size == 0:
        return np.zeros((0,), dtype=np.int64)
    order = np.lexsort((T.oq, -iou, T.ot))
    first = np.ones(order.size, dtype=bool)
    first[1:] = T.ot[order][1:] != T.ot[order][:-1]
    return order[first]


def aji_from_table(T, verbose=True):
    inst_t, inst_p = T.instance_lists()
    if inst_p.size == 0:
        raise ValueError("attempt to get an argmax of an empty sequence")      # np.argmax(axis=1) on [Nt, 0], :63
    inter = T.on.astype(np.float64)
    union = (T.area_t[T.ot] + T.area_q[T.oq] - T.on).astype(np.float64)          # :52-55
    iou = inter / (union + 1.0e-6)                                               # :60
    sel = _best_pred_per_true(T, iou)
    overall_inter = np.float64(T.on[sel].sum())                                  # :70-74: sums of integers, exact
    overall_union = np.float64((T.area_t[T.ot[sel]] + T.area_q[T.oq[sel]] - T.on[sel]).sum())
    overall_FP = np.float64((T.area_q[T.oq[sel]] - T.on[sel]).sum())
    overall_FN = np.float64((T.area_t[T.ot[sel]] - T.on[sel]).sum())
    unpaired_t = np.setdiff1d(inst_t, T.ot[sel])                                 # :80-81
    unpaired_p = np.setdiff1d(inst_p, T.oq[sel])
    less_pred = np.float64(T.area_t[unpaired_t].sum())                           # :86-91
    more_pred = np.float64(T.area_q[unpaired_p].sum())
    overall_union = overall_union + less_pred + more_pred
    aji_score = overall_inter / overall_union
    fm = overall_union - overall_inter
    res = (aji_score, overall_FP / fm, overall_FN / fm, less_pred / fm, more_pred / fm)
    if verbose:                                                                  # :95
        print('\t [ana_FP = {:.4f}, ana_FN = {:.4f}, ana_less = {:.4f}, ana_more = {:.4f}]'.format(*res[1:]))
    return res


def _dense_iou(T, inst_t, inst_p, eps):
    inter = np.zeros([inst_t.size, inst_p.size], dtype=np.float64)
    union = np.zeros([inst_t.size, inst_p.size], dtype=np.float64)
    inter[T.ot - 1, T.oq - 1] = T.on
    union[T.ot - 1, T.oq - 1] = T.area_t[T.ot] + T.area_q[T.oq] - T.on
    return inter, union, inter / (union + eps)


def aji_plus_from_table(T):
    from scipy.optimize import linear_sum_assignment   # the reference's own solver for the 1:1 pairing (:147)
    inst_t, inst_p = T.instance_lists()
    inter, union, iou = _dense_iou(T, inst_t, inst_p, 1.0e-6)
    pt, pp = linear_sum_assignment(-iou)
    piou = iou[pt, pp]
    pt, pp = pt[piou > 0.0], pp[piou > 0.0]
    overall_inter = (inter[pt, pp]).sum()
    overall_union = (union[pt, pp]).sum()
    overall_union = overall_union + np.float64(T.area_t[np.setdiff1d(inst_t, pt + 1)].sum()) \
        + np.float64(T.area_q[np.setdiff1d(inst_p, pp + 1)].sum())
    return overall_inter / overall_union


def pq_from_table(T, match_iou=0.5):
    assert match_iou >= 0.0, "Cant' be negative"
    inst_t, inst_p = T.instance_lists()
    iou = T.on.astype(np.float64) / (T.area_t[T.ot] + T.area_q[T.oq] - T.on).astype(np.float64)   # :232-234
    if match_iou >= 0.5:
        keep = iou > match_iou                     # rows are already in np.nonzero's (true, pred) order
        paired_true, paired_pred, paired_iou = T.ot[keep], T.oq[keep], iou[keep]
    else:
        from scipy.optimize import linear_sum_assignment
        dense = np.zeros([inst_t.size, inst_p.size], dtype=np.float64)
        dense[T.ot - 1, T.oq - 1] = iou
        pt, pp = linear_sum_assignment(-dense)
        paired_iou = dense[pt, pp]
        paired_true = list(pt[paired_iou > match_iou] + 1)
        paired_pred = list(pp[paired_iou > match_iou] + 1)
        paired_iou = paired_iou[paired_iou > match_iou]
    unpaired_true = list(np.setdiff1d(inst_t, np.asarray(paired_true, dtype=np.int64)))
    unpaired_pred = list(np.setdiff1d(inst_p, np.asarray(paired_pred, dtype=np.int64)))
    tp, fp, fn = len(paired_true), len(unpaired_pred), len(unpaired_true)
    dq = tp / (tp + 0.5 * fp + 0.5 * fn)           # ZeroDivisionError for two empty images, like the reference
    sq = paired_iou.sum() / (tp + 1.0e-6)
    return [dq, sq, dq * sq], [paired_true, paired_pred, unpaired_true, unpaired_pred]


def _dice2(T):
    if T.on.size == 0:
        return 2 * 0 / 0                           # the reference divides int 0 by int 0 here
    return 2 * np.uint64(T.on.sum()) / np.uint64((T.area_t[T.ot] + T.area_q[T.oq]).sum())


def dice1_from_table(T):
    inter = np.int64(T.on.sum())
    denom = np.int64(T.area_t[1:].sum() + T.area_q[1:].sum())
    return 2.0 * inter / denom


# ---- the reference's signatures -------------------------------------------------------------------------------
def get_fast_aji(true, pred):
    return aji_from_table(pair_table(true, pred))


def get_fast_aji_plus(true, pred):
    return aji_plus_from_table(pair_table(true, pred))


def get_fast_pq(true, pred, match_iou=0.5):
    assert match_iou >= 0.0, "Cant' be negative"
    return pq_from_table(pair_table(true, pred), match_iou)


def get_fast_dice_2(true, pred):
    T = pair_table(true, pred)
    T.instance_lists()
    return _dice2(T)


def get_dice_1(true, pred):
    return dice1_from_table(pair_table(true, pred))


def get_dice_2(true, pred):
    return _dice2(pair_table(true, pred))


def remap_label(pred, by_size=False):
    pred = np.asarray(pred)
    d = _to_device(pred)
    T = label_pairs_cuda(d, d)[0]                  # pairs (v, v): ids and sizes in one pass
    if T.area_t.size == 0 or T.area_t[0] == 0:
        raise ValueError("list.remove(x): x not in list")       # stats_utils.py:372 needs a background pixel
    ids = T.true_ids[1:]
    if ids.size == 0:
        return pred
    new = np.arange(1, ids.size + 1, dtype=np.int32)
    if by_size:
        rank = np.argsort(-T.area_t[ids], kind="stable")        # sorted(..., reverse=True) keeps ties in id order
        new = np.empty(ids.size, dtype=np.int32)
        new[rank] = np.arange(1, ids.size + 1, dtype=np.int32)
    L = _cabi.lib()
    dev = d.device
    out = torch.empty(d.shape, dtype=torch.int32, device=dev)
    s_ids = torch.from_numpy(ids.astype(np.int32)).to(dev)
    n_ids = torch.from_numpy(new).to(dev)
    check(L.cdnet_remap_labels(_ptr(d), 4 if d.dtype == torch.int32 else 8, _ptr(out), _ptr(s_ids), _ptr(n_ids),
                               int(ids.size), d.numel(), _stream()), "cdnet_remap_labels")
    return out[0].cpu().numpy()


def instance_metrics_cuda(true, pred, match_iou=0.5):
    """Batched: int32 / int64 tensors [B,H,W] on the GPU -> list of dicts with the scores test_dam.py:616-626
    reports (aji and its four error shares, dice, dq, sq, pq), one pair table per tile."""
    out = []
    for T in label_pairs_cuda(true, pred):
        aji = aji_from_table(T, verbose=False)
        (dq, sq, pq), _ = pq_from_table(T, match_iou)
        out.append({"aji": aji[0], "ana_FP": aji[1], "ana_FN": aji[2], "ana_less": aji[3], "ana_more": aji[4],
                    "dice": dice1_from_table(T), "dq": dq, "sq": sq, "pq": pq})
    return out


from .api import guard_public_functions as _guard  # noqa: E402  (device guard, see api._on_tensor_device)
_guard(globals())
