"""-m gpu parity tests of the instance-metrics path (stats_utils.py drop-ins): the CUDA pair table against
np.unique, the drop-in scores against goldens from the verbatim reference and against the oracle restatement."""
import contextlib
import io

import numpy as np
import pytest

from conftest import load_golden, to_dev
from metrics_numpy_pairs import pair_arrays
from test_metrics_host import cases, inputs, check_against_golden, same, _edge_inputs, _outcome, _same_outcome

@pytest.fixture
def M(kernel_api):
    from cdnet_b200 import metrics
    return metrics


def _sorted_pairs(keys, counts):
    o = np.argsort(keys)
    return keys[o], np.asarray(counts, dtype=np.int64)[o]


@pytest.mark.parametrize("H,W,n,dtype", [(64, 80, 6, np.int32), (97, 143, 14, np.int64), (1, 37, 1, np.int32),
                                         (333, 517, 150, np.int32), (1000, 1000, 700, np.int64)])
def test_pair_table_vs_numpy(M, H, W, n, dtype):
    import torch
    from cdnet_b200 import synth
    true, pred = synth.metric_pair(500 + H, H, W, n, 1)
    t = to_dev(M, torch.from_numpy(true.astype(dtype))[None])
    p = to_dev(M, torch.from_numpy(pred.astype(dtype))[None])
    T = M.label_pairs_cuda(t, p)[0]
    rk, rc = _sorted_pairs(*pair_arrays(true, pred))
    got = (T.t.astype(np.uint64) << np.uint64(32)) | T.q.astype(np.uint64)
    assert np.array_equal(got, rk) and np.array_equal(T.n, rc)
    assert int(T.n.sum()) == H * W


def test_pair_table_batch_and_small_cap(M):
    """Several tiles in one launch; a cap far too small takes the overflow -> retry path and still is exact."""
    import torch
    from cdnet_b200 import synth
    pairs = [synth.metric_pair(600 + i, 120, 136, 15, i % 2) for i in range(5)]
    t = to_dev(M, torch.from_numpy(np.stack([a for a, _ in pairs])))
    p = to_dev(M, torch.from_numpy(np.stack([b for _, b in pairs])))
    for cap in (None, 3):
        tabs = M.label_pairs_cuda(t, p, cap=cap)
        for (a, b), T in zip(pairs, tabs):
            rk, rc = _sorted_pairs(*pair_arrays(a, b))
            got = (T.t.astype(np.uint64) << np.uint64(32)) | T.q.astype(np.uint64)
            assert np.array_equal(got, rk) and np.array_equal(T.n, rc)


def test_metrics_golden(M):
    z, cs = cases()
    for c in cs:
        true, pred = inputs(c)
        with contextlib.redirect_stdout(io.StringIO()):
            aji = M.get_fast_aji(true, pred)
        check_against_golden(z, c["name"], aji, M.get_fast_aji_plus(true, pred),
                             {mi: M.get_fast_pq(true, pred, mi) for mi in (0.5, 0.3)},
                             M.get_dice_1(true, pred), M.get_fast_dice_2(true, pred))
        raw = (true.astype(np.int64) * 3 + (true > 0) * 5).astype(np.int32)
        assert np.array_equal(M.remap_label(raw), z[c["name"] + "_remap"])
        assert np.array_equal(M.remap_label(raw, by_size=True), z[c["name"] + "_remap_by_size"])


@pytest.mark.parametrize("dtype", [np.uint8, np.int32, np.int64, np.uint16])
def test_metrics_vs_oracle_dtypes(M, dtype):
    from cdnet_b200 import synth
    from oracle import restate_metrics as O
    true, pred = synth.metric_pair(77, 150, 170, 30, 1)
    true, pred = true.astype(dtype), pred.astype(dtype)
    with contextlib.redirect_stdout(io.StringIO()):
        assert same(M.get_fast_aji(true, pred), O.get_fast_aji(true, pred))
    assert same(M.get_fast_pq(true, pred), O.get_fast_pq(true, pred))
    assert same(M.get_fast_pq(true, pred, 0.7), O.get_fast_pq(true, pred, 0.7))
    assert same(M.get_dice_1(true, pred), O.get_dice_1(true.astype(np.int64), pred.astype(np.int64)))
    assert same(M.get_dice_2(true, pred), O.get_dice_2(true, pred))
    assert same(M.get_fast_aji_plus(true, pred), O.get_fast_aji_plus(true, pred))


def test_metrics_edge_cases(M):
    from oracle import restate_metrics as O
    for name, (t, p) in _edge_inputs().items():
        for i, (fo, fp) in enumerate([(lambda: O.get_fast_aji(t, p), lambda: M.get_fast_aji(t, p)),
                                      (lambda: O.get_fast_pq(t, p), lambda: M.get_fast_pq(t, p)),
                                      (lambda: O.get_dice_1(t, p), lambda: M.get_dice_1(t, p)),
                                      (lambda: O.get_dice_2(t, p), lambda: M.get_dice_2(t, p)),
                                      (lambda: O.get_fast_dice_2(t, p), lambda: M.get_fast_dice_2(t, p))]):
            a, b = _outcome(fo), _outcome(fp)
            assert _same_outcome(a, b), (name, i, a, b)
    with pytest.raises(ValueError):
        M.get_dice_1(np.full((4, 4), -1, np.int32), np.zeros((4, 4), np.int32))
    with pytest.raises(ValueError):                      # stats_utils.py:372: no background id to remove
        M.remap_label(np.ones((4, 4), np.int32))
    z = np.zeros((4, 4), np.int32)
    assert M.remap_label(z) is z or np.array_equal(M.remap_label(z), z)


def test_metrics_after_postprocessing(M, cuda_api):
    """Downstream witness: post-process a tile on the GPU, score it against the synthetic
    ground truth on the GPU; the scores equal the oracle's scores of the verbatim reference's labels."""
    import torch
    from cdnet_b200 import synth
    from oracle import restate_metrics as O
    z, meta = load_golden("p_256")
    d = synth.postproc_inputs(meta["seed"], meta["H"], meta["W"], meta["n_target"])
    lab = cuda_api.dam_postprocess(d["prob"].copy(), d["point"], d["dcm"], meta["direction_classes"], meta["min_area"],
                                   meta["radius"], 0)
    gt = synth.contiguous_ids(d["ids"])
    pred = M.remap_label(lab)
    ref_pred = O.remap_label(z["dam_pp0_labels"])
    assert np.array_equal(pred, ref_pred)
    res = M.instance_metrics_cuda(to_dev(M, torch.from_numpy(gt)[None]), to_dev(M, torch.from_numpy(pred)[None]))[0]
    with contextlib.redirect_stdout(io.StringIO()):
        ref_aji = O.get_fast_aji(gt, ref_pred)
    assert same([res["aji"], res["ana_FP"], res["ana_FN"], res["ana_less"], res["ana_more"]], list(ref_aji))
    assert same(res["dice"], O.get_dice_1(gt, ref_pred))
    assert same([res["dq"], res["sq"], res["pq"]], O.get_fast_pq(gt, ref_pred)[0])
    assert res["aji"] > 0.3      # the synthetic prediction is a sensible segmentation of the synthetic truth
