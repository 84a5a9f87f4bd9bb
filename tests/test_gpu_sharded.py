"""-m gpu: whole-slide row partition with the CUDA backend, all ranks simulated on one GPU (SimComm):
the sharded result must equal the unsharded single-GPU result (and the oracle) bit for bit."""
import os
import sys

import numpy as np
import pytest

from conftest import to_dev

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def _slide(seed, H, W, n, n_maps):
    from test_sharded_gloo import _slide as mk
    return mk(seed, H, W, n, n_maps)


@pytest.mark.parametrize("G", [2, 3, 4, 7])
@pytest.mark.parametrize("n_maps", [8, 1])
def test_sharded_equals_single_gpu(kernel_api, G, n_maps):
    import torch
    from cdnet_b200 import sharded
    H, W = 612, 524
    dcm, prob, point = _slide(41, H, W, 230, n_maps)
    single, status = kernel_api.dam_postprocess_cuda(to_dev(kernel_api, torch.from_numpy(dcm)[None]), to_dev(kernel_api, torch.from_numpy(prob)[None]),
                                                   to_dev(kernel_api, torch.from_numpy(point)[None]), 9, 20, 2, 0)
    assert int(status.sum()) == 0
    single = single[0].cpu().numpy()
    parts = sharded.row_partition(H, G)
    shards = [dict(dcm=dcm[:, a:b].copy(), prob=prob[:, a:b].copy(), point=point[:, a:b].copy()) for a, b in parts]
    be = sharded.CudaBackend()
    outs = sharded.postprocess_slide(shards, sharded.SimComm(G), H, W, be, 9, 20, 2)
    got = np.concatenate([o.cpu().numpy() for o in outs], axis=0)
    assert got.dtype == single.dtype
    assert np.array_equal(got, single), int((got != single).sum())


def test_sharded_equals_oracle(kernel_api):
    from cdnet_b200 import sharded
    from oracle import restate as O
    H, W = 300, 280
    dcm, prob, point = _slide(42, H, W, 60, 8)
    ref = O.dam_postprocess(prob.copy(), point, dcm, 9, 20, 2, 0, literal=False)["pred_labeled"]
    parts = sharded.row_partition(H, 3)
    shards = [dict(dcm=dcm[:, a:b].copy(), prob=prob[:, a:b].copy(), point=point[:, a:b].copy()) for a, b in parts]
    outs = sharded.postprocess_slide(shards, sharded.SimComm(3), H, W, sharded.CudaBackend(), 9, 20, 2)
    got = np.concatenate([o.cpu().numpy() for o in outs], axis=0)
    assert np.array_equal(got, ref)


def test_single_map_variant_vs_oracle(kernel_api):
    """n_maps = 1 (test_dam.py:499-502) through the unsharded C ABI"""
    import torch
    from oracle import restate as O
    from cdnet_b200 import synth
    d = synth.postproc_inputs(43, 200, 240, 30)
    dcm = d["dcm"][:1].copy()
    ref = O.dam_postprocess(d["prob"].copy(), d["point"], dcm, 9, 20, 2, 0, literal=False)["pred_labeled"]
    out, _ = kernel_api.dam_postprocess_cuda(to_dev(kernel_api, torch.from_numpy(dcm)[None]), to_dev(kernel_api, torch.from_numpy(d["prob"])[None]),
                                           to_dev(kernel_api, torch.from_numpy(d["point"])[None]), 9, 20, 2, 0)
    assert np.array_equal(out[0].cpu().numpy(), ref)


def test_sharded_wide_slide(kernel_api):
    """W > 16384 (what a 40 000-wide slide uses): shards == single GPU"""
    import torch
    from cdnet_b200 import sharded, synth
    base = synth.postproc_inputs(78, 30, 1040, 40)
    d = {k: np.ascontiguousarray(np.concatenate([base[k]] * 16, axis=-1)) for k in ("dcm", "prob", "point")}
    H, W = 30, d["dcm"].shape[-1]
    dcm = d["dcm"][:1].copy()
    single, _ = kernel_api.dam_postprocess_cuda(to_dev(kernel_api, torch.from_numpy(dcm)[None]), to_dev(kernel_api, torch.from_numpy(d["prob"])[None]),
                                              to_dev(kernel_api, torch.from_numpy(d["point"])[None]), 9, 20, 2, 0)
    parts = sharded.row_partition(H, 3)
    shards = [dict(dcm=dcm[:, a:b].copy(), prob=d["prob"][:, a:b].copy(), point=d["point"][:, a:b].copy()) for a, b in parts]
    outs = sharded.postprocess_slide(shards, sharded.SimComm(3), H, W, sharded.CudaBackend(), 9, 20, 2)
    got = np.concatenate([o.cpu().numpy() for o in outs], axis=0)
    assert np.array_equal(got, single[0].cpu().numpy())


@pytest.mark.parametrize("seed,H,W,G", [(51, 64, 40, 8), (52, 97, 132, 6), (54, 120, 64, 7)])
def test_sharded_thin_shards(kernel_api, seed, H, W, G):
    """thin shards: components and holes that cross several seams (device-side seam rounds)"""
    import torch
    from cdnet_b200 import sharded
    from test_sharded_gloo import _slide as mk
    rng = np.random.default_rng(seed)
    dcm, prob, point = mk(seed, H, W, max(4, H * W // 900), 8)
    yy, xx = np.mgrid[0:H, 0:W]
    snake = ((yy // 3) % 2 == 0) & (xx > 2) & (xx < W - 3)
    link = ((yy % 6) == 3) & (xx >= W - 6) & (xx < W - 3) | ((yy % 6) == 0) & (xx > 2) & (xx <= 5) & (yy > 0)
    prob[1][(snake | link) & (rng.random((H, W)) < 0.97)] += np.float32(3.0)
    single, _ = kernel_api.dam_postprocess_cuda(to_dev(kernel_api, torch.from_numpy(dcm)[None]), to_dev(kernel_api, torch.from_numpy(prob)[None]),
                                              to_dev(kernel_api, torch.from_numpy(point)[None]), 9, 20, 2, 0)
    parts = sharded.row_partition(H, G)
    shards = [dict(dcm=dcm[:, a:b].copy(), prob=prob[:, a:b].copy(), point=point[:, a:b].copy()) for a, b in parts]
    outs = sharded.postprocess_slide(shards, sharded.SimComm(G), H, W, sharded.CudaBackend(), 9, 20, 2)
    got = np.concatenate([o.cpu().numpy() for o in outs], axis=0)
    assert np.array_equal(got, single[0].cpu().numpy()), int((got != single[0].cpu().numpy()).sum())


def _plain_slide(seed, H, W, n):
    from cdnet_b200 import synth
    d = synth.postproc_inputs(seed, H, W, n)
    return d["dcm"].copy(), d["prob"], d["point"]


@pytest.mark.parametrize("G,n_maps,overlap", [(2, 8, 48), (3, 8, 40), (4, 1, 64), (3, 8, 100)])
def test_sharded_watershed_equals_single_gpu(kernel_api, G, n_maps, overlap):
    """postproc = 1 (process(): EDT, markers, watershed) on a row-sharded slide: own rows + overlap rows per rank,
    slide-global marker ids agreed between the ranks == the unsharded call, bit for bit"""
    import torch
    from cdnet_b200 import sharded
    H, W = 412, 356
    dcm, prob, point = _plain_slide(61 + G, H, W, 120)
    dcm = dcm[:n_maps].copy()
    single, status = kernel_api.dam_postprocess_cuda(to_dev(kernel_api, torch.from_numpy(dcm)[None]),
                                                     to_dev(kernel_api, torch.from_numpy(prob)[None]),
                                                     to_dev(kernel_api, torch.from_numpy(point)[None]), 9, 20, 2, 1)
    assert int(status.sum() & 0xff) == 0
    single = single[0].cpu().numpy()
    assert single.max() > 20
    parts = sharded.row_partition(H, G)
    shards = [dict(dcm=dcm[:, a:b].copy(), prob=prob[:, a:b].copy(), point=point[:, a:b].copy()) for a, b in parts]
    outs = sharded.postprocess_slide(shards, sharded.SimComm(G), H, W, sharded.CudaBackend(), 9, 20, 2, postproc=1,
                                     overlap=overlap)
    got = np.concatenate([o.cpu().numpy() for o in outs], axis=0)
    assert got.dtype == single.dtype == np.int32
    assert np.array_equal(got, single), int((got != single).sum())


def test_sharded_watershed_vs_oracle(kernel_api):
    from cdnet_b200 import sharded
    from oracle import restate as O
    H, W = 260, 240
    dcm, prob, point = _plain_slide(66, H, W, 50)
    ref = O.dam_postprocess(prob.copy(), point, dcm, 9, 20, 2, 1, literal=False)["pred_labeled"]
    parts = sharded.row_partition(H, 3)
    shards = [dict(dcm=dcm[:, a:b].copy(), prob=prob[:, a:b].copy(), point=point[:, a:b].copy()) for a, b in parts]
    outs = sharded.postprocess_slide(shards, sharded.SimComm(3), H, W, sharded.CudaBackend(), 9, 20, 2, postproc=1,
                                     overlap=48)
    got = np.concatenate([o.cpu().numpy() for o in outs], axis=0)
    assert np.array_equal(got, ref), int((got != ref).sum())


def test_sharded_watershed_overflow_is_reported(kernel_api):
    """a structure taller than the overlap (the seam-straddling bar of _slide) cannot be normalised per shard:
    the call must raise, not return different labels"""
    from cdnet_b200 import sharded
    H, W = 300, 280
    dcm, prob, point = _slide(42, H, W, 60, 8)
    parts = sharded.row_partition(H, 3)
    shards = [dict(dcm=dcm[:, a:b].copy(), prob=prob[:, a:b].copy(), point=point[:, a:b].copy()) for a, b in parts]
    with pytest.raises(RuntimeError, match="overlap"):
        sharded.postprocess_slide(shards, sharded.SimComm(3), H, W, sharded.CudaBackend(), 9, 20, 2, postproc=1,
                                  overlap=32)
    with pytest.raises(ValueError, match="overlap"):
        sharded.postprocess_slide(shards, sharded.SimComm(3), H, W, sharded.CudaBackend(), 9, 20, 2, postproc=1,
                                  overlap=128)


def test_shard_ws_process_entry(kernel_api):
    """cdnet_shard_ws_process through the C ABI: labels == the oracle's process(); marker_rowmax == the per-row maximum
    of label4(markers) BEFORE remove_small_objects (postproc_other.py:44 vs :46); the overflow bit for a component that
    reaches from an outer overlap row into the own rows, and only then"""
    import torch
    from scipy import ndimage as ndi
    from cdnet_b200 import _cabi, synth
    from oracle import restate as O
    be = __import__("cdnet_b200.sharded", fromlist=["CudaBackend"]).CudaBackend()
    d = synth.postproc_inputs(71, 180, 200, 40)
    pred = (np.argmax(d["prob"], axis=0) == 1).astype(np.uint8)
    ref, parts = O.process(pred.astype(np.float64).copy(), min_size=10, return_parts=True, literal=False)
    marker0 = ndi.binary_erosion(ndi.binary_fill_holes(parts["dist"] > 125), iterations=1)
    rowmax_ref = O.label4(marker0).max(axis=1)
    labels, rowmax, status = be.ws_process(to_dev(kernel_api, torch.from_numpy(pred)), 0, 180, 10)
    assert int(status[0]) & 0xff == 0
    assert np.array_equal(labels.cpu().numpy(), ref)
    assert np.array_equal(rowmax.cpu().numpy(), rowmax_ref)
    # own rows [60, 120): nuclei are small, nothing reaches row 0 or row 179 from there
    _, _, status = be.ws_process(to_dev(kernel_api, torch.from_numpy(pred)), 60, 120, 10)
    assert not int(status[0]) & _cabi.S_SHARD_OVERFLOW
    for col, rows in ((20, slice(0, 70)), (150, slice(110, 180))):   # a bar from the top / bottom edge into the own rows
        p2 = pred.copy()
        p2[rows, col:col + 3] = 1
        _, _, status = be.ws_process(to_dev(kernel_api, torch.from_numpy(p2)), 60, 120, 10)
        assert int(status[0]) & _cabi.S_SHARD_OVERFLOW
    p2 = pred.copy()
    p2[0:55, 20:23] = 1      # ends before the own rows: fine
    p2[58:125, 90:93] = 1    # crosses the own rows but stays inside the tile: fine
    _, _, status = be.ws_process(to_dev(kernel_api, torch.from_numpy(p2)), 60, 120, 10)
    assert not int(status[0]) & _cabi.S_SHARD_OVERFLOW


def test_shard_ws_relabel_entry(kernel_api):
    """cdnet_shard_ws_relabel: owned ids by arithmetic, adopted ids through the table, err for an id without owner"""
    import torch
    be = __import__("cdnet_b200.sharded", fromlist=["CudaBackend"]).CudaBackend()
    rng = np.random.default_rng(5)
    lab = rng.integers(0, 40, size=(37, 53)).astype(np.int32)
    above, owned, off = 10, 20, 1000
    lut = np.zeros(64, np.int32)
    lut[1:11] = 500 + np.arange(10)       # adopted from above
    lut[31:40] = 2000 + np.arange(9)      # adopted from below
    want = np.where((lab > above) & (lab <= above + owned), lab - above + off, lut[lab]).astype(np.int32)
    dev = lambda a: to_dev(kernel_api, torch.from_numpy(a))
    out = be.empty(lab.shape, "int32")
    err = be.ws_relabel(dev(lab), dev(np.array([above, owned, off], np.int32)), dev(lut), out)
    assert int(err[0]) == 0 and np.array_equal(out.cpu().numpy(), want)
    lut[35] = 0
    err = be.ws_relabel(dev(lab), dev(np.array([above, owned, off], np.int32)), dev(lut), out)
    assert int(err[0]) == 1


@pytest.mark.parametrize("G,radius,overlap,out_dtype", [(2, 0, 60, "int64"), (5, 1, 44, "int32"), (3, 2, 52, "int64")])
def test_sharded_watershed_radius_and_dtype(kernel_api, G, radius, overlap, out_dtype):
    """postproc = 1 with every dilation radius of the reference and both output types; a crowded slide (touching nuclei
    that the watershed splits, small markers dropped next to the seams) so that marker ids with gaps cross the seams"""
    import torch
    from cdnet_b200 import sharded, synth
    H, W = 330, 300
    d = synth.postproc_inputs(90 + G, H, W, 160)
    dcm, prob, point = d["dcm"].copy(), d["prob"], d["point"]
    single, _ = kernel_api.dam_postprocess_cuda(to_dev(kernel_api, torch.from_numpy(dcm)[None]),
                                                to_dev(kernel_api, torch.from_numpy(prob)[None]),
                                                to_dev(kernel_api, torch.from_numpy(point)[None]), 9, 20, radius, 1,
                                                out_dtype=getattr(torch, out_dtype))
    single = single[0].cpu().numpy()
    ids = np.unique(single)
    assert len(ids) > 30 and ids.max() > len(ids)          # gaps in the numbering (markers dropped as too small)
    parts = sharded.row_partition(H, G)
    shards = [dict(dcm=dcm[:, a:b].copy(), prob=prob[:, a:b].copy(), point=point[:, a:b].copy()) for a, b in parts]
    outs = sharded.postprocess_slide(shards, sharded.SimComm(G), H, W, sharded.CudaBackend(), 9, 20, radius,
                                     out_dtype=out_dtype, postproc=1, overlap=overlap)
    got = np.concatenate([o.cpu().numpy() for o in outs], axis=0)
    assert got.dtype == single.dtype
    assert np.array_equal(got, single), int((got != single).sum())
