"""CPU tests of the instance-metrics path: the oracle restatement against goldens made
by the verbatim reference (stats_utils.py) and against the reference itself where it is mounted; the product's host
epilogue (cdnet_b200/metrics.py) on pair tables from a numpy stand-in of the CUDA reduction."""
import contextlib
import importlib.util
import io
import os

import numpy as np
import pytest

from conftest import load_golden
from metrics_numpy_pairs import pair_arrays


def bits(x):
    return np.asarray(x, dtype=np.float64).view(np.int64)


def same(a, b):
    if isinstance(a, (list, tuple)):
        return len(a) == len(b) and all(same(x, y) for x, y in zip(a, b))
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape:
        return False
    if a.dtype.kind == "f" or b.dtype.kind == "f":
        return np.array_equal(bits(a), bits(b))
    return np.array_equal(a, b)


def cases():
    z, meta = load_golden("metrics")
    return z, meta["cases"]


def inputs(c):
    from cdnet_b200 import synth
    true, pred = synth.metric_pair(c["seed"], c["H"], c["W"], c["n_target"], c["mode"])
    assert synth.digest(true, pred) == c["digest"], "synthetic inputs differ from the goldens'"
    return true, pred


def check_against_golden(z, name, aji, aji_plus, pq, dice1, dice2):
    assert same(np.asarray(aji, dtype=np.float64), z[name + "_aji"]), name
    assert same(np.float64(aji_plus), z[name + "_aji_plus"]), name
    for mi, res in pq.items():
        tag = name + "_pq%02d" % int(mi * 10)
        (dq, sq, p), (pt, pp, ut, up) = res
        assert same(np.asarray([dq, sq, p], dtype=np.float64), z[tag]), tag
        assert np.array_equal(np.asarray(pt, dtype=np.int64), z[tag + "_paired_true"]), tag
        assert np.array_equal(np.asarray(pp, dtype=np.int64), z[tag + "_paired_pred"]), tag
        assert np.array_equal(np.asarray(ut, dtype=np.int64), z[tag + "_unpaired_true"]), tag
        assert np.array_equal(np.asarray(up, dtype=np.int64), z[tag + "_unpaired_pred"]), tag
    assert same(np.float64(dice1), z[name + "_dice1"]), name
    assert same(np.float64(dice2), z[name + "_dice2"]), name


def test_oracle_metrics_golden():
    from oracle import restate_metrics as M
    z, cs = cases()
    for c in cs:
        true, pred = inputs(c)
        check_against_golden(z, c["name"], M.get_fast_aji(true, pred), M.get_fast_aji_plus(true, pred),
                             {mi: M.get_fast_pq(true, pred, mi) for mi in (0.5, 0.3)},
                             M.get_dice_1(true, pred), M.get_fast_dice_2(true, pred))
        raw = (true.astype(np.int64) * 3 + (true > 0) * 5).astype(np.int32)
        assert np.array_equal(M.remap_label(raw), z[c["name"] + "_remap"])
        assert np.array_equal(M.remap_label(raw, by_size=True), z[c["name"] + "_remap_by_size"])


def test_host_epilogue_golden():
    """cdnet_b200.metrics on pair tables (numpy stand-in for the kernel) == the verbatim reference's scores."""
    from cdnet_b200 import metrics as P
    z, cs = cases()
    for c in cs:
        true, pred = inputs(c)
        T = P.PairTable(*pair_arrays(true, pred), n_pixels=true.size)
        check_against_golden(z, c["name"], P.aji_from_table(T, verbose=False), P.aji_plus_from_table(T),
                             {mi: P.pq_from_table(T, mi) for mi in (0.5, 0.3)}, P.dice1_from_table(T), P._dice2(T))


def _edge_inputs():
    z = np.zeros((16, 16), np.int32)
    o = z.copy()
    o[2:5, 2:5] = 1
    two = o.copy()
    two[8:12, 8:12] = 2
    gap = o.copy()
    gap[8:12, 8:12] = 3          # non-contiguous ids
    full = np.ones((8, 8), np.int32)
    return {"empty-pred": (o, z), "empty-true": (z, o), "both-empty": (z, z), "identical": (two, two),
            "gap-true": (gap, two), "gap-pred": (two, gap), "no-background": (full, full)}


def _outcome(f):
    try:
        with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
            return ("ok", f())
    except Exception as e:  # noqa: BLE001 -- the error TYPE is the thing compared
        return ("exc", type(e).__name__)


def _same_outcome(a, b):
    if a[0] != b[0]:
        return False
    if a[0] == "exc":
        return a[1] == b[1]
    return same(a[1], b[1]) or str(a[1]) == str(b[1])   # nan compares by text


def test_host_epilogue_edge_cases_vs_oracle():
    from cdnet_b200 import metrics as P
    from oracle import restate_metrics as M
    for name, (t, p) in _edge_inputs().items():
        T = lambda: P.PairTable(*pair_arrays(t, p), n_pixels=t.size)  # noqa: E731
        pairs = [(lambda: M.get_fast_aji(t, p), lambda: P.aji_from_table(T(), verbose=False)),
                 (lambda: M.get_fast_pq(t, p), lambda: P.pq_from_table(T())),
                 (lambda: M.get_dice_1(t, p), lambda: P.dice1_from_table(T())),
                 (lambda: M.get_dice_2(t, p), lambda: P._dice2(T()))]
        for i, (fo, fp) in enumerate(pairs):
            a, b = _outcome(fo), _outcome(fp)
            assert _same_outcome(a, b), (name, i, a, b)


@pytest.mark.skipif(not os.path.exists("/root/reference/stats_utils.py"), reason="reference not mounted")
def test_oracle_metrics_vs_verbatim_reference():
    from cdnet_b200 import synth
    from oracle import restate_metrics as M
    spec = importlib.util.spec_from_file_location("ref_stats_utils", "/root/reference/stats_utils.py")
    R = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(R)
    for seed, H, W, n, mode in [(41, 72, 90, 8, 1), (42, 130, 117, 25, 0), (43, 160, 160, 30, 1)]:
        t, p = synth.metric_pair(seed, H, W, n, mode)
        with contextlib.redirect_stdout(io.StringIO()):
            ra = R.get_fast_aji(t, p)
        assert same(ra, M.get_fast_aji(t, p))
        assert same(R.get_fast_aji_plus(t, p), M.get_fast_aji_plus(t, p))
        for mi in (0.5, 0.3, 0.7):
            assert same(R.get_fast_pq(t, p, mi), M.get_fast_pq(t, p, mi))
        assert same(R.get_fast_dice_2(t, p), M.get_fast_dice_2(t, p))
        assert same(R.get_dice_1(t, p), M.get_dice_1(t, p))
        if H * W <= 72 * 90:
            assert same(R.get_dice_2(t, p), M.get_dice_2(t, p))
        raw = (t.astype(np.int64) * 3).astype(np.int32)
        assert same(R.remap_label(raw), M.remap_label(raw)) and same(R.remap_label(raw, True), M.remap_label(raw, True))
    # error behaviour on the edge cases where the reference itself is well defined (contiguous ids)
    for name, (t, p) in _edge_inputs().items():
        if name.startswith("gap") or name == "no-background":
            continue
        for fr, fo in [(lambda: R.get_fast_aji(t, p), lambda: M.get_fast_aji(t, p)),
                       (lambda: R.get_fast_pq(t, p), lambda: M.get_fast_pq(t, p)),
                       (lambda: R.get_dice_1(t, p), lambda: M.get_dice_1(t, p)),
                       (lambda: R.get_fast_dice_2(t, p), lambda: M.get_fast_dice_2(t, p))]:
            a, b = _outcome(fr), _outcome(fo)
            assert _same_outcome(a, b), (name, a, b)
