"""-m gpu parity tests of the inference post-processing path: CUDA (through the C ABI) vs the golden
vectors generated from the verbatim reference, and vs the oracle restatement on seeded inputs."""
import numpy as np
import pytest

from conftest import load_golden, to_dev


def _inputs(meta):
    from cdnet_b200 import synth
    d = synth.postproc_inputs(meta["seed"], meta["H"], meta["W"], meta["n_target"])
    assert synth.digest(d["dcm"], d["prob"], d["point"]) == meta["digest"], "synthetic inputs differ from the goldens'"
    return d


def test_ddm_golden(kernel_api):
    z, meta = load_golden("ddm")
    for key, cls in meta["cases"]:
        out = kernel_api.generate_dd_map(z[key + "_in"], cls)
        ref = z[key + "_out"]
        assert out.dtype == np.float32 and out.shape == ref.shape
        assert np.array_equal(out, ref, equal_nan=True), key


def test_circshift_golden(kernel_api):
    z, meta = load_golden("ddm")
    for i, (d, s1, s2) in enumerate(meta["circshift"]):
        out = kernel_api.circshift(z["cs_in"], d, s1, s2)
        assert out.dtype == z["cs_in"].dtype
        assert np.array_equal(out, z["cs_%d" % i]), (d, s1, s2)


@pytest.mark.parametrize("name", ["p_96x128", "p_200x150", "p_256", "p_333x517", "p_1000"])
@pytest.mark.parametrize("postproc", [0, 1])
def test_dam_postprocess_golden(kernel_api, name, postproc):
    z, meta = load_golden(name)
    d = _inputs(meta)
    prob = d["prob"].copy()
    lab = kernel_api.dam_postprocess(prob, d["point"], d["dcm"], meta["direction_classes"], meta["min_area"],
                                   meta["radius"], postproc)
    ref = z["dam_pp%d_labels" % postproc]
    assert str(lab.dtype) == str(z["dam_pp%d_dtype" % postproc])
    assert lab.shape == ref.shape
    assert np.array_equal(lab, ref), "%d differing pixels" % int((lab != ref).sum())


@pytest.mark.parametrize("name", ["p_96x128", "p_256", "p_333x517"])
@pytest.mark.parametrize("postproc", [0, 1])
def test_plain_postprocess_golden(kernel_api, name, postproc):
    z, meta = load_golden(name)
    d = _inputs(meta)
    lab = kernel_api.plain_postprocess(d["prob"].copy(), meta["min_area"], meta["radius"], postproc)
    ref = z["plain_pp%d_labels" % postproc]
    assert str(lab.dtype) == str(z["plain_pp%d_dtype" % postproc])
    assert np.array_equal(lab, ref), "%d differing pixels" % int((lab != ref).sum())


def test_process_golden(kernel_api):
    z, meta = load_golden("process")
    for key, shape in meta["cases"]:
        mask = np.unpackbits(z[key + "_in"])[:shape[0] * shape[1]].reshape(shape).astype(bool)
        for ms in (5, 10):
            src = mask.astype(np.uint8) * 255
            out = kernel_api.process(src, "modelName", min_size=ms)
            assert out.dtype == np.int32
            assert set(np.unique(src)) <= {0, 1}, "process() must binarise its input in place"
            assert np.array_equal(out, z["%s_ms%d" % (key, ms)]), (key, ms)
        out = kernel_api.process(mask.astype(np.uint8) * 255, "unet", min_size=10)
        assert np.array_equal(out, z["%s_unet" % key]), key


def test_dam_mutates_prob_like_reference(kernel_api):
    from oracle import restate as O
    from cdnet_b200 import synth
    d = synth.postproc_inputs(3, 120, 90, 10)
    p_ref = d["prob"].copy()
    O.dam_postprocess(p_ref, d["point"], d["dcm"], 9, 20, 2, 0)
    p_gpu = d["prob"].copy()
    kernel_api.dam_postprocess(p_gpu, d["point"], d["dcm"], 9, 20, 2, 0)
    assert np.array_equal(p_ref.view(np.uint32), p_gpu.view(np.uint32))


def test_dam_constant_direction_map_asserts(kernel_api):
    from cdnet_b200 import synth
    d = synth.postproc_inputs(3, 64, 64, 4)
    d["dcm"][3] = 0  # constant map -> NaN DDM -> the reference's assert fires (test_dam.py:535)
    with pytest.raises(AssertionError):
        kernel_api.dam_postprocess(d["prob"].copy(), d["point"], d["dcm"], 9, 20, 2, 0)


@pytest.mark.parametrize("seed,H,W,n", [(201, 77, 131, 12), (202, 301, 299, 70), (203, 512, 384, 140)])
def test_dam_postprocess_vs_oracle(kernel_api, seed, H, W, n):
    """fresh seeds / ragged sizes: CUDA vs the oracle restatement run on the box's host"""
    from oracle import restate as O
    from cdnet_b200 import synth
    d = synth.postproc_inputs(seed, H, W, n)
    for cls in (9,):
        for pp in (0, 1):
            ref = O.dam_postprocess(d["prob"].copy(), d["point"], d["dcm"], cls, 20, 2, pp, literal=False)
            lab = kernel_api.dam_postprocess(d["prob"].copy(), d["point"], d["dcm"], cls, 20, 2, pp)
            assert np.array_equal(lab, ref["pred_labeled"]), (seed, pp, int((lab != ref["pred_labeled"]).sum()))


def _outcome(fn):
    try:
        return ("ok", fn())
    except (AssertionError, ValueError) as e:
        return (type(e).__name__, None)


def test_dam_postprocess_edge_cases(kernel_api):
    """degenerate tiles (synth.postproc_edge_cases; the oracle is pinned to the verbatim reference on the same
    cases in tests/test_oracle_vs_reference.py): same labels, or the same exception type as the reference"""
    from oracle import restate as O
    from cdnet_b200 import synth
    for name, c in synth.postproc_edge_cases():
        for pp in (0, 1):
            ref = _outcome(lambda: O.dam_postprocess(c["prob"].copy(), c["point"], c["dcm"], 9, 20, 2, pp,
                                                     literal=False)["pred_labeled"])
            got = _outcome(lambda: kernel_api.dam_postprocess(c["prob"].copy(), c["point"], c["dcm"], 9, 20, 2, pp))
            assert got[0] == ref[0], (name, pp, got[0], ref[0])
            if ref[1] is not None:
                assert got[1].dtype == ref[1].dtype and np.array_equal(got[1], ref[1]), (name, pp)
            ref = _outcome(lambda: O.plain_postprocess(c["prob"].copy(), 20, 2, pp, literal=False)["pred_labeled"])
            got = _outcome(lambda: kernel_api.plain_postprocess(c["prob"].copy(), 20, 2, pp))
            assert got[0] == ref[0], (name, pp, "plain", got[0], ref[0])
            if ref[1] is not None:
                assert got[1].dtype == ref[1].dtype and np.array_equal(got[1], ref[1]), (name, pp, "plain")


def test_process_all_foreground_raises(kernel_api):
    """postproc_other.py:18-19: `nuc_list.remove(0)` raises ValueError when the mask has no background pixel"""
    import torch
    full = np.full((24, 40), 255, np.uint8)
    with pytest.raises(ValueError):
        kernel_api.process(full.copy(), "modelName")
    assert kernel_api.process(full.copy(), "unet").max() == 1  # the no-watershed head has no such list
    # device-resident batch: the status bit marks exactly the tile without background
    m = np.ones((3, 24, 40), np.uint8)
    m[0, 3, 4] = 0
    m[2, :, 20:] = 0
    _, st = kernel_api.process_cuda(to_dev(kernel_api, torch.from_numpy(m)), 10, True, return_status=True)
    assert [int(v) & 16 for v in st.cpu().numpy()] == [0, 16, 0]


def test_primitives_vs_scipy(kernel_api):
    from scipy import ndimage as ndi
    from oracle import restate as O
    rng = np.random.default_rng(5)
    # ragged widths take the scalar kernels, W % 4 == 0 the 4-pixel ones, W > 1024 crosses init row chunks,
    # W > 16384 takes the row-tree merge instead of the shared-memory strips
    for H, W, p in ((1, 1, 0.5), (1, 40, 0.6), (37, 1, 0.6), (65, 129, 0.55), (200, 333, 0.62), (128, 128, 0.95),
                    (64, 64, 0.0), (31, 33, 1.0), (7, 8, 0.6), (3, 12, 0.7), (50, 64, 0.58), (5, 2052, 0.6),
                    (9, 4100, 0.62), (4, 16500, 0.6), (3, 17001, 0.6)):
        m = rng.random((H, W)) < p
        assert np.array_equal(kernel_api.label(m, connectivity=1), ndi.label(m)[0]), (H, W, "label4")
        assert np.array_equal(kernel_api.label(m), O.label8(m)), (H, W, "label8")
        assert np.array_equal(kernel_api.binary_fill_holes(m), ndi.binary_fill_holes(m)), (H, W, "fill")
        assert np.array_equal(kernel_api.remove_small_objects(m, 7), O.remove_small_objects(m, 7)), (H, W, "rso")
        if not m.all():
            assert np.array_equal(kernel_api.distance_transform_edt(m), ndi.distance_transform_edt(m)), (H, W, "edt")
        lab = ndi.label(m)[0]
        for r in (1, 2):
            assert np.array_equal(kernel_api.dilation(lab, radius=r), O.dilate(lab, O.disk(r))), (H, W, "dil", r)
        assert np.array_equal(kernel_api.remove_small_objects(lab, 5), O.remove_small_objects(lab, 5))


def test_batched_equals_single(kernel_api):
    import torch
    from cdnet_b200 import synth
    tiles = [synth.postproc_inputs(300 + i, 96, 160, 12) for i in range(3)]
    dcm = to_dev(kernel_api, torch.from_numpy(np.stack([t["dcm"] for t in tiles])))
    prob = to_dev(kernel_api, torch.from_numpy(np.stack([t["prob"] for t in tiles])))
    point = to_dev(kernel_api, torch.from_numpy(np.stack([t["point"] for t in tiles])))
    for pp in (0, 1):
        out, status = kernel_api.dam_postprocess_cuda(dcm, prob.clone(), point, 9, 20, 2, pp)
        assert int(status.abs().sum()) == 0
        for i, t in enumerate(tiles):
            single = kernel_api.dam_postprocess(t["prob"].copy(), t["point"], t["dcm"], 9, 20, 2, pp)
            assert np.array_equal(out[i].cpu().numpy(), single)


def test_wide_tile_pipeline_vs_oracle(kernel_api):
    """W > 16384: the whole-slide fallback path (row-chunk init + row-tree merge) through the full pipeline"""
    from oracle import restate as O
    from cdnet_b200 import synth
    base = synth.postproc_inputs(77, 24, 1100, 40)
    rep = 16
    d = {k: np.ascontiguousarray(np.concatenate([base[k]] * rep, axis=-1)) for k in ("dcm", "prob", "point")}
    assert d["dcm"].shape[-1] == 17600
    for pp in (0, 1):
        ref = O.dam_postprocess(d["prob"].copy(), d["point"], d["dcm"], 9, 20, 2, pp, literal=False)["pred_labeled"]
        got = kernel_api.dam_postprocess(d["prob"].copy(), d["point"], d["dcm"], 9, 20, 2, pp)
        assert np.array_equal(got, ref), (pp, int((got != ref).sum()))


def test_dcm_voting2_vs_oracle(kernel_api):
    from oracle import restate as O
    rng = np.random.default_rng(11)
    dm = rng.integers(0, 9, size=(57, 83, 8)).astype(np.uint8)
    got = kernel_api.DcmVoting2(dm)
    assert got.dtype == np.int64 and np.array_equal(got, O.dcm_voting2(dm))


def test_direction_argmax_handoff(kernel_api):
    """device-resident hand-off == the reference's softmax / argmax done by torch (test_dam.py:984-1013)"""
    import torch
    g = torch.Generator(device="cpu").manual_seed(3)
    mask_logits = torch.randn((2, 3, 40, 48), generator=g)
    dir_logits = torch.randn((2, 9, 40, 48), generator=g)
    prob, cls = kernel_api.direction_argmax_cuda(to_dev(kernel_api, mask_logits), to_dev(kernel_api, dir_logits))
    for i in range(2):
        p = torch.softmax(to_dev(kernel_api, mask_logits[i]), dim=0)
        d = torch.softmax(to_dev(kernel_api, dir_logits[i]), dim=0)
        d[0] = d[0] * p[0]
        top2 = torch.topk(d, 2, dim=0).values
        clear = (top2[0] - top2[1]) > 1e-6   # an exact tie-break is a property of torch's kernels, not ours
        assert torch.equal(cls[i].long()[clear], torch.argmax(d, dim=0)[clear])
        assert torch.allclose(prob[i], p, rtol=1e-6, atol=1e-7)


def _tta_inputs(seed, B, H, W, C):
    import torch
    g = torch.Generator().manual_seed(seed)
    shapes = [(H, W)] * 4 + [(W, H)] * 4
    ml = [torch.randn((B, 3) + s, generator=g) * 3 for s in shapes]
    pt = [torch.randn((B, 1) + s, generator=g) for s in shapes]
    dl = [torch.randn((B, C) + s, generator=g) * 3 for s in shapes]
    return ml, pt, dl


@pytest.mark.parametrize("B,H,W,C", [(1, 21, 34, 9), (2, 64, 96, 9), (1, 33, 31, 17), (1, 1, 1, 5), (1, 70, 5, 9)])
def test_tta_merge_vs_oracle(kernel_api, B, H, W, C):
    """fused TTA hand-off (test_dam.py:299-450, :983-1013) vs the restatement (pinned to the verbatim reference in
    tests/test_oracle_vs_reference.py).  Probabilities: 1e-5 relative (the softmax's expf is the device's, the
    reference's is torch's); point map: bit-exact; direction classes: exact wherever the top-2 margin exceeds
    the float noise."""
    from oracle import restate as O
    ml, pt, dl = _tta_inputs(B * 100 + H, B, H, W, C)
    prob, point, dcm = kernel_api.tta_merge_cuda([to_dev(kernel_api, t) for t in ml], [to_dev(kernel_api, t) for t in pt],
                                                 [to_dev(kernel_api, t) for t in dl])
    assert prob.dtype == kernel_api.torch.float32 and dcm.dtype == kernel_api.torch.uint8
    assert tuple(prob.shape) == (B, 3, H, W) and tuple(point.shape) == (B, 1, H, W) and tuple(dcm.shape) == (B, 8, H, W)
    for b in range(B):
        rp, rq, rd = O.tta_merge([t[b].numpy() for t in ml], [t[b].numpy() for t in pt], [t[b].numpy() for t in dl])
        assert np.allclose(prob[b].cpu().numpy(), rp, rtol=1e-5, atol=1e-7)
        assert np.array_equal(point[b].cpu().numpy().view(np.uint32), rq.view(np.uint32))
        got = dcm[b].cpu().numpy().astype(np.int64)
        for v in range(8):
            p, _, _ = O.variant_probmaps(ml[v][b].numpy(), pt[v][b].numpy(), dl[v][b].numpy())
            z = dl[v][b].numpy().astype(np.float64)
            q = np.exp(z - z.max(axis=0)) / np.exp(z - z.max(axis=0)).sum(axis=0)
            q[0] *= p[0]
            top = np.sort(q, axis=0)
            clear = O.tta_variant_to_original(((top[-1] - top[-2]) > 1e-6)[None], v)[0]
            assert np.array_equal(got[v][clear], rd[v][clear]), (b, v)
            assert clear.mean() > 0.99


def test_tta_merge_feeds_postprocess(kernel_api):
    """hand-off -> dam_postprocess_cuda without leaving the device == the same two steps through the oracle"""
    from oracle import restate as O
    from cdnet_b200 import synth
    torch = kernel_api.torch
    d = synth.postproc_inputs(88, 96, 80, 9)
    H, W = 96, 80
    # logits whose softmax / argmax reproduce the synthetic tile in every variant's frame
    def to_variant(a, v):
        a = np.asarray(a)
        if v & 4:
            a = np.rot90(a, k=1, axes=(1, 2))
        if v & 2:
            a = np.flip(a, 1)
        if v & 1:
            a = np.flip(a, 2)
        return np.ascontiguousarray(a)
    ml, pt, dl = [], [], []
    for v in range(8):
        ml.append(torch.from_numpy(to_variant(np.log(d["prob"] + 1e-6), v))[None])
        pt.append(torch.from_numpy(to_variant(d["point"], v))[None])
        onehot = (np.arange(9)[:, None, None] == d["dcm"][v][None]).astype(np.float32) * 12.0
        dl.append(torch.from_numpy(to_variant(onehot, v))[None])
    for v in range(8):  # to_variant inverts tta_variant_to_original
        assert np.array_equal(O.tta_variant_to_original(to_variant(d["point"], v), v), d["point"])
    prob, point, dcm = kernel_api.tta_merge_cuda([to_dev(kernel_api, t) for t in ml], [to_dev(kernel_api, t) for t in pt],
                                                 [to_dev(kernel_api, t) for t in dl])
    assert np.array_equal(dcm[0].cpu().numpy(), d["dcm"])
    lab, status = kernel_api.dam_postprocess_cuda(dcm, prob, point, 9, 20, 2, 0)
    assert int(status[0]) == 0
    ref = O.dam_postprocess(prob[0].cpu().numpy().copy(), point[0].cpu().numpy(), dcm[0].cpu().numpy(), 9, 20, 2, 0,
                            literal=False)["pred_labeled"]
    assert np.array_equal(lab[0].cpu().numpy(), ref)
