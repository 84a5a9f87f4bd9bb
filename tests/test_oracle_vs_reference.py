"""CPU, build container only: the oracle restatement against the reference executed VERBATIM from
/root/reference (oracle/ref_loader.py) on fresh seeded inputs.  Skipped where the reference tree is
not mounted (the GPU box)."""
import numpy as np
import pytest

from cdnet_b200 import synth
from oracle import ref_loader, restate as O

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not mounted")


@pytest.fixture(scope="module")
def ref():
    return ref_loader.load()


def test_generate_dd_map(ref):
    rng = np.random.default_rng(9)
    for cls in (5, 9, 17):
        x = rng.integers(0, cls, size=(41, 67)).astype(np.uint8)
        x[rng.integers(0, 2, size=x.shape) == 0] = 0
        assert np.array_equal(ref.generate_dd_map(x, cls), O.generate_dd_map(x, cls), equal_nan=True)


def test_quantiser(ref):
    ang = np.concatenate([np.linspace(-180, 180, 1441), [22.5, 22.500002, 67.5, -157.5, 157.5, 180.0, -180.0]])
    ang = ang.astype(np.float32)
    n = ref.DTOffsetConfig.num_classes
    a_ref, i_ref = ref.DTOffsetHelper.align_angle(ang.copy(), num_classes=n)
    a_o, i_o = O.align_angle(ang.copy(), n)
    assert np.array_equal(i_ref, i_o) and np.array_equal(a_ref, a_o)
    v_ref = ref.DTOffsetHelper.angle_to_vector(ang.copy(), num_classes=n)
    assert np.array_equal(v_ref, O.angle_to_vector(ang.copy(), n))
    assert np.array_equal(ref.DTOffsetHelper.vector_to_label(v_ref, num_classes=n), O.vector_to_label(v_ref, n))
    # Appendix B.6: the double quantisation is idempotent
    assert np.array_equal(O.vector_to_label(O.angle_to_vector(ang, n), n), i_o)
    assert np.array_equal(ref.Sobel.kernel(ksize=11).numpy().reshape(2, 11, 11), O.sobel_kernels(11))


def test_postprocess_and_process(ref):
    d = synth.postproc_inputs(777, 150, 170, 20)
    for pp in (0, 1):
        r = ref.dam_postprocess(d["prob"].copy(), d["point"], d["dcm"], 9, 20, 2, pp)
        o = O.dam_postprocess(d["prob"].copy(), d["point"], d["dcm"], 9, 20, 2, pp)
        assert r["pred_labeled"].dtype == o["pred_labeled"].dtype
        assert np.array_equal(r["pred_labeled"], o["pred_labeled"])
    m = (synth.instance_map(778, 120, 140, 18) > 0).astype(np.uint8) * 255
    assert np.array_equal(ref.process(m.copy(), "modelName", min_size=5), O.process(m.copy(), "modelName", min_size=5))


def _outcome(fn):
    try:
        return ("ok", fn())
    except (AssertionError, ValueError) as e:
        return (type(e).__name__, None)


def test_postprocess_edge_cases(ref):
    """degenerate tiles: the restatement returns what the verbatim reference returns, or raises the same
    exception type (AssertionError test_dam.py:535, ValueError postproc_other.py:19)"""
    import warnings
    for name, c in synth.postproc_edge_cases():
        for pp in (0, 1):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                r = _outcome(lambda: ref.dam_postprocess(c["prob"].copy(), c["point"], c["dcm"], 9, 20, 2, pp)["pred_labeled"])
                o = _outcome(lambda: O.dam_postprocess(c["prob"].copy(), c["point"], c["dcm"], 9, 20, 2, pp)["pred_labeled"])
                o2 = _outcome(lambda: O.dam_postprocess(c["prob"].copy(), c["point"], c["dcm"], 9, 20, 2, pp,
                                                        literal=False)["pred_labeled"])
            for got in (o, o2):
                assert got[0] == r[0], (name, pp, got[0], r[0])
                if r[1] is not None:
                    assert got[1].dtype == r[1].dtype and np.array_equal(got[1], r[1]), (name, pp)
            r = _outcome(lambda: ref.plain_postprocess(c["prob"].copy(), 20, 2, pp)["pred_labeled"])
            o = _outcome(lambda: O.plain_postprocess(c["prob"].copy(), 20, 2, pp, literal=False)["pred_labeled"])
            assert o[0] == r[0], (name, pp, "plain", o[0], r[0])
            if r[1] is not None:
                assert o[1].dtype == r[1].dtype and np.array_equal(o[1], r[1]), (name, pp, "plain")


def test_label_encoding(ref):
    lab = synth.as_uint8_label(synth.instance_map(779, 110, 130, 12))
    res = ref.LabelEncoding(3, 1, 1)((None, None, lab))
    n = ref.DTOffsetConfig.num_classes
    for literal in (True, False):
        tern, point, direction = O.label_encoding(lab, num_classes=n, literal=literal)
        assert np.array_equal(np.asarray(res[2]), tern)
        assert np.array_equal(res[3].view(np.uint16), point.view(np.uint16))
        assert np.array_equal(res[4], direction)


def test_label_encoding_edge_cases(ref):
    """degenerate label images, [H,W,3] and [H,W]: verbatim reference == restatement (both forms)"""
    n = ref.DTOffsetConfig.num_classes
    for name, ids in synth.label_edge_cases():
        for lab in (np.repeat(ids[:, :, None], 3, axis=2), ids):
            res = ref.LabelEncoding(3, 1, 1)((None, None, lab.copy()))
            for literal in (True, False):
                tern, point, direction = O.label_encoding(lab.copy(), num_classes=n, literal=literal)
                assert np.array_equal(np.asarray(res[2]), tern), (name, lab.ndim, literal)
                assert np.array_equal(res[3].view(np.uint16), point.view(np.uint16)), (name, lab.ndim, literal)
                assert np.array_equal(res[4], direction), (name, lab.ndim, literal)


def test_dcm_voting2(ref):
    rng = np.random.default_rng(3)
    dm = rng.integers(0, 9, size=(33, 47, 8)).astype(np.uint8)
    assert np.array_equal(ref.DcmVoting2(dm), O.dcm_voting2(dm))


def test_training_consumers(ref):
    """train_util_dam.py:123-142 and my_transforms.LabelEncoding (no direction), verbatim vs restatement"""
    import torch
    rng = np.random.default_rng(17)
    for B, H, W, C in ((2, 19, 23, 9), (3, 8, 8, 17)):
        d = rng.integers(0, C, size=(B, H, W)).astype(np.int64)
        t = rng.integers(0, 3, size=(B, H, W)).astype(np.int64)
        d[B - 1] = 3
        r = ref.direction_one_hot(torch.from_numpy(d.copy()), torch.from_numpy(t.copy()), C).numpy()
        o = O.direction_one_hot(d, t, C)
        assert r.dtype == o.dtype and np.array_equal(r, o)
    lab = synth.as_uint8_label(synth.instance_map(781, 90, 110, 14))
    binary = np.repeat(((lab[:, :, 0] > 0) * 255).astype(np.uint8)[:, :, None], 3, axis=2)
    for img in (lab, binary, lab[:, :, 0].copy()):
        for out_c in ((3, 1) if img.ndim == 3 else (3,)):
            r = np.asarray(ref.LabelEncodingPlain(out_c, 1, 0)((None, None, img.copy()))[2])
            assert np.array_equal(r, O.label_encoding_plain(img.copy(), out_c)), (img.shape, out_c)


def test_label_encoding_plain_with_direction(ref):
    """my_transforms.LabelEncoding(3, 1, do_direction=1) verbatim (peak_local_max from the shim) vs the restatement,
    literal (full-canvas per-instance loops) and windowed"""
    for seed, H, W, n in ((31, 72, 90, 9), (32, 130, 110, 25)):
        ids = synth.instance_map(seed, H, W, n)
        for lab in (synth.as_uint8_label(ids), np.repeat(((ids > 0) * 255).astype(np.uint8)[:, :, None], 3, axis=2)):
            r = ref.LabelEncodingPlain(3, 1, 1)((None, None, lab.copy()))
            for literal in (True, False):
                o = O.label_encoding_plain_direction(lab.copy(), 3, 8, literal=literal)
                assert np.array_equal(np.asarray(r[2]), o[0])
                assert np.array_equal(r[3].view(np.uint16), o[1].view(np.uint16))
                assert np.array_equal(r[4], o[2])


def test_tta_merge(ref):
    """the TTA block (test_dam.py:314-450) and get_probmaps (:930-1034) executed verbatim with stand-in model /
    image objects vs the restatement: point maps and direction classes exact, probabilities to float32 rounding
    (torch's CPU softmax vs numpy's exp)"""
    import torch
    g = torch.Generator().manual_seed(5)
    for H, W, C in ((21, 34, 9), (16, 16, 17)):
        shapes = [(H, W)] * 4 + [(W, H)] * 4
        ml = [torch.randn((3,) + s, generator=g) * 3 for s in shapes]
        pt = [torch.randn((1,) + s, generator=g) for s in shapes]
        dl = [torch.randn((C,) + s, generator=g) * 3 for s in shapes]
        rp, rq, rd = ref.tta_merge(ml, pt, dl)
        op, oq, od = O.tta_merge([m.numpy() for m in ml], [p.numpy() for p in pt], [d.numpy() for d in dl])
        assert rp.dtype == op.dtype == np.float32 and rp.shape == op.shape
        assert np.allclose(rp, op, rtol=1e-6, atol=1e-7)
        assert np.array_equal(rq.view(np.uint32), oq.view(np.uint32))
        assert (rd != od).mean() < 1e-3  # exact float ties aside
        for v in range(8):  # closed-form index maps used by the kernel
            a = np.arange(shapes[v][0] * shapes[v][1]).reshape((1,) + shapes[v])
            b = O.tta_variant_to_original(a, v)[0]
            yy, xx = np.mgrid[0:H, 0:W]
            for y, x in ((0, 0), (H - 1, 0), (0, W - 1), (H - 1, W - 1), (3, 7), (11, 2)):
                r, c = O.tta_source_index(v, y, x, H, W)
                assert b[y, x] == a[0, r, c]


def test_sobel_drop_in(ref):
    from cdnet_b200.training import Sobel
    for k in (3, 5, 11, 15):
        a, b = ref.Sobel.kernel(ksize=k), Sobel.kernel(ksize=k)
        assert a.dtype == b.dtype and a.shape == b.shape and bool((a == b).all()), k
    assert Sobel.kernel() is Sobel.kernel(11)


def test_label_encoding_out_c_1(ref):
    """out_c != 3 (my_transforms_direction.py:721-739) with and without direction targets: verbatim reference ==
    restatement (both forms), same exception types on the inputs the reference cannot index"""
    n = ref.DTOffsetConfig.num_classes
    lab = synth.as_uint8_label(synth.instance_map(779, 70, 90, 6))
    binary = np.repeat(((lab[:, :, 0] > 0) * 255).astype(np.uint8)[:, :, None], 3, axis=2)
    shifted = binary.copy()
    shifted[:, :, 1] = np.roll(binary[:, :, 0], 5, axis=1)
    const1 = lab.copy()
    const1[:, :, 1] = 7
    cases = [lab, binary, shifted, const1, lab[:, :, 0].copy()]
    cases += [np.repeat(ids[:, :, None], 3, axis=2) for _, ids in synth.label_edge_cases()]
    for i, img in enumerate(cases):
        for dd in (1, 0):
            try:
                r = ("ok", ref.LabelEncoding(1, 1, dd)((None, None, img.copy())))
            except Exception as e:  # noqa: BLE001 -- the point is the exception TYPE
                r = (type(e).__name__, None)
            for literal in (True, False):
                try:
                    o = ("ok", O.label_encoding(img.copy(), out_c=1, do_direction=dd, num_classes=n, literal=literal))
                except Exception as e:  # noqa: BLE001
                    o = (type(e).__name__, None)
                assert o[0] == r[0], (i, dd, literal, o[0], r[0])
                if r[1] is None:
                    continue
                assert len(r[1]) == (5 if dd else 3)
                assert np.array_equal(np.asarray(r[1][2]), o[1][0]), (i, dd, literal)
                if dd:
                    assert np.array_equal(r[1][3].view(np.uint16), o[1][1].view(np.uint16)), (i, dd, literal)
                    assert np.array_equal(r[1][4], o[1][2]), (i, dd, literal)


def test_postprocess_voting_first(ref):
    """the block's `voting_firt` switch flipped in the text that is exec-ed (test_dam.py:471): DcmVoting2, one DDM"""
    d = synth.postproc_inputs(780, 130, 150, 16)
    for pp in (0, 1):
        r = ref.dam_postprocess(d["prob"].copy(), d["point"], d["dcm"], 9, 20, 2, pp, voting_first=True)
        o = O.dam_postprocess(d["prob"].copy(), d["point"], d["dcm"], 9, 20, 2, pp, voting_first=True)
        assert r["pred_labeled"].dtype == o["pred_labeled"].dtype
        assert np.array_equal(r["pred_labeled"], o["pred_labeled"])
        plain = O.dam_postprocess(d["prob"].copy(), d["point"], d["dcm"], 9, 20, 2, pp)
        assert not np.array_equal(plain["ddm_mean"], o["ddm_mean"])  # the switch does change the map


def test_postprocess_unet_model_mode(ref):
    d = synth.postproc_inputs(783, 110, 140, 14)
    r = ref.plain_postprocess(d["prob"].copy(), 20, 2, 1, model_name="unet")["pred_labeled"]
    o = O.plain_postprocess(d["prob"].copy(), 20, 2, 1, model_name="unet", literal=False)["pred_labeled"]
    assert r.dtype == o.dtype and np.array_equal(r, o)
    r = ref.dam_postprocess(d["prob"].copy(), d["point"], d["dcm"], 9, 20, 2, 1, model_name="unet")["pred_labeled"]
    o = O.dam_postprocess(d["prob"].copy(), d["point"], d["dcm"], 9, 20, 2, 1, model_name="unet")["pred_labeled"]
    assert r.dtype == o.dtype and np.array_equal(r, o)
