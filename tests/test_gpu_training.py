"""Kernel parity tests of the training-side operators (csrc/training.cu): the stand-alone direction quantiser
(DTOffsetHelper), the direction one-hot block and my_transforms.LabelEncoding without direction, against the
goldens generated from the verbatim reference (tests/golden/training.npz) and against the oracle restatement.
[cuda] on the B200 through the C ABI, [simt] on the host under the SIMT emulator."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from cdnet_b200 import synth


@pytest.fixture
def T(kernel_api):
    from cdnet_b200 import training
    return training


def _inputs():
    z, meta = load_golden("training")
    d = synth.training_inputs()
    assert synth.digest(*[d[k] for k in sorted(d)]) == meta["digest"], "synthetic inputs differ from the goldens'"
    return z, meta, d


def _same(a, b):
    return a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


def test_quantiser_golden(T):
    z, meta, d = _inputs()
    H = T.DTOffsetHelper
    for tag in ("32", "64"):
        a = d["angle" + tag]
        for n in (8, 4):
            s, i = H.align_angle(a.copy(), num_classes=n)
            assert _same(s, z["align%d_np%s_s" % (n, tag)]) and _same(i, z["align%d_np%s_i" % (n, tag)]), (tag, n)
            s, i = H.align_angle(torch.from_numpy(a.copy()), num_classes=n, return_tensor=True)
            assert isinstance(s, torch.Tensor) and isinstance(i, torch.Tensor)
            assert _same(s.numpy(), z["align%d_pt%s_s" % (n, tag)]) and _same(i.numpy(), z["align%d_pt%s_i" % (n, tag)])
            assert _same(H.angle_to_vector(a.copy(), num_classes=n), z["a2v%d_np%s" % (n, tag)]), (tag, n)
            v = H.angle_to_vector(torch.from_numpy(a.copy()), num_classes=n, return_tensor=True)
            assert _same(v.numpy(), z["a2v%d_pt%s" % (n, tag)]), (tag, n)
        # bin centres are far from every bin edge: exact.  Random vectors: the class may differ only where the
        # angle is within 1e-4 degrees of an edge (host libm atan2 vs the device's)
        assert _same(H.vector_to_label(z["a2v8_np" + tag].copy(), num_classes=8), z["v2l8_roundtrip" + tag])
        got, ref = H.vector_to_label(d["vec" + tag].copy(), num_classes=8), z["v2l8_np" + tag]
        assert got.dtype == ref.dtype
        for y, x in np.argwhere(got != ref):
            v = d["vec" + tag][y, x].astype(np.float64)
            ang = np.degrees(np.arctan2(v[0], v[1]))
            e = abs((ang + 180.0 - 22.5) % 45.0)
            assert min(e, 45.0 - e) < 1e-4, (tag, y, x, ang)


def test_quantiser_16_and_32_vs_oracle(T):
    from oracle import restate as O
    d = synth.training_inputs()
    for n in (16, 32):
        for tag in ("32", "64"):
            a = d["angle" + tag]
            s, i = T.DTOffsetHelper.align_angle(a.copy(), num_classes=n)
            so, io = O.align_angle(a.copy(), n)
            assert _same(s, so) and _same(i, io), (n, tag)
            assert _same(T.DTOffsetHelper.angle_to_vector(a.copy(), num_classes=n), O.angle_to_vector(a.copy(), n))


def test_quantiser_ragged_and_unaligned(T):
    """n % 4 != 0 (scalar tail after the 128-bit quads) and a base pointer that is not 16-byte aligned"""
    from oracle import restate as O
    d = synth.training_inputs()
    flat = d["angle32"].reshape(-1)
    for a in (flat[:1001], flat[:3], flat[1:1002], flat[3:]):
        t = torch.from_numpy(a.copy())
        if a is not flat[:1001] and a is not flat[:3]:
            t = torch.from_numpy(flat.copy())[flat.size - a.size:]  # a view: data pointer offset by 4 or 12 bytes
            a = t.numpy()
        so, io = O.align_angle(a.copy(), 8)
        s, i = T.DTOffsetHelper.align_angle(a.copy(), num_classes=8)
        assert _same(s, so) and _same(i, io)
        st, it = T.DTOffsetHelper.align_angle(t, num_classes=8, return_tensor=True)
        assert _same(it.numpy(), io) and _same(st.numpy(), so.astype(np.float32))
        assert _same(T.DTOffsetHelper.angle_to_vector(a.copy(), num_classes=8), O.angle_to_vector(a.copy(), 8))
        v = O.angle_to_vector(a.copy(), 8)
        assert _same(T.DTOffsetHelper.vector_to_label(v.copy(), num_classes=8), O.vector_to_label(v.copy(), 8))
        vt = torch.from_numpy(np.concatenate([np.zeros((1, 2)), v]).astype(np.float32))[1:]  # offset by 8 bytes
        assert _same(T.DTOffsetHelper.vector_to_label(vt, num_classes=8, return_tensor=True).numpy(),
                     O.vector_to_label(v.copy(), 8))


def test_label_to_vector_golden(T):
    z, meta, d = _inputs()
    lab = torch.from_numpy(d["labels17"])
    for C in (4, 5, 8, 9, 16, 17, 32):
        out = T.DTOffsetHelper.label_to_vector(lab, num_classes=C)
        assert out.device == lab.device and _same(out.numpy(), z["l2v%d" % C]), C
    for dt in (torch.uint8, torch.int32, torch.float32):
        out = T.DTOffsetHelper.label_to_vector(lab.to(dt), num_classes=17)
        assert _same(out.numpy(), z["l2v17"]), dt
    with pytest.raises(KeyError):
        T.DTOffsetHelper.label_to_vector(lab, num_classes=7)
    with pytest.raises(AssertionError):
        T.DTOffsetHelper.label_to_vector(d["labels17"], num_classes=9)  # numpy in: the reference asserts a tensor


def test_angle_to_direction_label_vs_oracle(T):
    from oracle import restate as O
    d = synth.training_inputs()
    a = d["angle32"]
    rng = np.random.default_rng(2)
    seg = rng.integers(-1, 2, size=a.shape)
    dist = rng.uniform(0, 10, size=a.shape)
    extra = rng.random(a.shape) < 0.1
    ref = O.align_angle(a, 8)[1].copy()
    ref[dist > T.DTOffsetConfig.max_distance] = 8
    ref[(seg == -1) | extra] = -1
    got = T.DTOffsetHelper.angle_to_direction_label(a.copy(), seg_label_map=seg, distance_map=dist, num_classes=8,
                                                    extra_ignore_mask=extra)
    assert _same(got, ref)


def test_direction_one_hot_golden(T):
    z, meta, d = _inputs()
    dirs, tern = torch.from_numpy(d["onehot_dir"]), torch.from_numpy(d["onehot_target"])
    out = T.direction_one_hot(dirs, tern, 9)
    assert out.dtype == torch.float32 and _same(out.numpy(), z["onehot9"])
    assert _same(T.direction_one_hot(dirs, tern.to(torch.uint8), 9).numpy(), z["onehot9"])
    # tile order matters only through target[0] (train_util_dam.py:139): swapping tiles 0 and 1 changes the mask
    from oracle import restate as O
    perm = [1, 0, 2, 3]
    assert _same(T.direction_one_hot(dirs[perm], tern[perm], 9).numpy(),
                 O.direction_one_hot(d["onehot_dir"][perm], d["onehot_target"][perm], 9))


@pytest.mark.parametrize("B,H,W,C", [(1, 1, 1, 9), (2, 7, 9, 9), (3, 23, 37, 17), (2, 16, 64, 5)])
def test_direction_one_hot_vs_oracle(T, B, H, W, C):
    from oracle import restate as O
    rng = np.random.default_rng(B * 1000 + H)
    dirs = rng.integers(0, C, size=(B, H, W)).astype(np.int64)
    tern = rng.integers(0, 3, size=(B, H, W)).astype(np.int64)
    assert _same(T.direction_one_hot(torch.from_numpy(dirs), torch.from_numpy(tern), C).numpy(),
                 O.direction_one_hot(dirs, tern, C))
    if H * W > 1:
        dirs[B - 1, 0, 0] = C  # out of range -> the reference's indexing raises
        with pytest.raises(IndexError):
            T.direction_one_hot(torch.from_numpy(dirs), torch.from_numpy(tern), C)


def test_label_encoding_plain_golden(T):
    z, meta, d = _inputs()
    cases = dict(synth.label_edge_cases())
    for name in meta["plain"]:
        ids = cases[name]
        lab = np.repeat(ids[:, :, None], 3, axis=2)
        lab[:, :, 1] = np.where(ids > 0, 0, 255)[..., ::-1]
        for out_c in (3, 1):
            res = T.LabelEncoding(out_c, 1, 0)(("img", "weight", lab.copy()))
            assert len(res) == 3 and res[0] == "img" and res[1] == "weight"
            got = np.asarray(res[2])
            assert got.dtype == np.uint8 and np.array_equal(got, z["plain_%s_c%d" % (name, out_c)]), (name, out_c)
        # a 2-D label image takes the same branches for out_c == 3 and cannot be indexed for out_c != 3
        assert np.array_equal(np.asarray(T.LabelEncoding(3, 1, 0)((None, None, ids.copy()))[2]), z["plain_%s_c3" % name])
        with pytest.raises(IndexError):
            T.LabelEncoding(1, 1, 0)((None, None, ids.copy()))
    with pytest.raises(NotImplementedError):
        T.LabelEncoding(2, 1, 1)  # do_direction = 1 is built for out_c = 3 only


@pytest.mark.parametrize("seed,H,W,n", [(61, 97, 143, 14), (62, 256, 200, 60), (63, 1, 9, 1), (64, 33, 1, 1)])
def test_label_encoding_plain_vs_oracle(T, seed, H, W, n):
    from oracle import restate as O
    ids = synth.instance_map(seed, max(H, 16), max(W, 16), n)[:H, :W]
    lab = np.ascontiguousarray(synth.as_uint8_label(ids))
    for out_c in (3, 1):
        assert np.array_equal(np.asarray(T.LabelEncoding(out_c, 1, 0)((None, None, lab.copy()))[2]),
                              O.label_encoding_plain(lab.copy(), out_c)), out_c
    binary = np.repeat(((ids > 0) * 255).astype(np.uint8)[:, :, None], 3, axis=2)
    for out_c in (3, 1):
        assert np.array_equal(np.asarray(T.LabelEncoding(out_c, 1, 0)((None, None, binary.copy()))[2]),
                              O.label_encoding_plain(binary.copy(), out_c)), out_c


@pytest.mark.parametrize("seed,H,W,n,classes", [(3, 64, 64, 5, 8), (4, 96, 120, 14, 8), (5, 150, 130, 30, 16), (6, 200, 256, 60, 8)])
def test_label_encoding_plain_direction_vs_oracle(T, seed, H, W, n, classes):
    """my_transforms.LabelEncoding(3, 1, do_direction=1) (my_transforms.py:763-836): centres from the nucleus's own
    distance transform; instance ids and {0,255} labels.  The restatement is pinned to the verbatim reference in
    tests/test_oracle_vs_reference.py (tie order among equal EDT maxima: unpinned, raster-first on both sides)."""
    from oracle import restate as O
    from cdnet_b200 import synth
    ids = synth.instance_map(seed, H, W, n)
    binary = np.repeat(((ids > 0) * 255).astype(np.uint8)[:, :, None], 3, axis=2)
    for lab in (synth.as_uint8_label(ids), binary):
        ref = O.label_encoding_plain_direction(lab.copy(), 3, classes, literal=False)
        res = T.LabelEncoding(3, 1, 1, num_classes=classes)((None, None, lab.copy()))
        assert len(res) == 5
        assert np.array_equal(np.asarray(res[2]), ref[0])
        assert res[3].dtype == np.float16 and np.array_equal(res[3].view(np.uint16), ref[1].view(np.uint16))
        assert res[4].dtype == np.int64 and np.array_equal(res[4], ref[2]), int((res[4] != ref[2]).sum())


def test_label_encoding_plain_direction_scope(T):
    with pytest.raises(NotImplementedError):
        T.LabelEncoding(2, 1, 1)


def test_label_encoding_plain_direction_golden(T):
    """goldens from the verbatim my_transforms.LabelEncoding(3, 1, 1)"""
    z, meta = load_golden("plaindir")
    for c in meta["cases"]:
        ids = synth.instance_map(c["seed"], c["H"], c["W"], c["n"])
        labs = {"inst": synth.as_uint8_label(ids), "bin": np.repeat(((ids > 0) * 255).astype(np.uint8)[:, :, None], 3, axis=2)}
        for kind, lab in labs.items():
            res = T.LabelEncoding(3, 1, 1, num_classes=8)((None, None, lab.copy()))
            key = "%s_%s_" % (c["name"], kind)
            assert np.array_equal(np.asarray(res[2]), z[key + "ternary"]), key
            assert np.array_equal(res[3].view(np.uint16), z[key + "point"].view(np.uint16)), key
            assert np.array_equal(res[4], z[key + "direction"].astype(np.int64)), key
