"""Kernel parity tests added while widening round 1 (degenerate-input sweeps, the all-foreground error path, the
fused TTA hand-off): same two backends as the other kernel tests -- [cuda] on the B200 through the C ABI, [simt] on
the host under the SIMT emulator.  Kept in a file that sorts after the others so that `pytest -x` reaches the long
established parity tests first."""
import numpy as np
import pytest

from conftest import load_golden, to_dev
from test_gpu_targets import _check_direction


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


def test_label_encoding_degenerate_sweep(kernel_api):
    """synth.label_edge_cases (pinned reference -> oracle in tests/test_oracle_vs_reference.py), as [H,W,3] and
    [H,W] label images, 8 and 16 direction classes"""
    from oracle import restate as O
    from cdnet_b200 import synth
    for name, ids in synth.label_edge_cases():
        for lab in (np.repeat(ids[:, :, None], 3, axis=2), ids):
            for n in (8, 16):
                ref = O.label_encoding(lab.copy(), num_classes=n, literal=False)
                res = kernel_api.LabelEncoding(3, 1, 1, num_classes=n)((None, None, lab.copy()))
                assert np.array_equal(np.asarray(res[2]), ref[0]), (name, lab.ndim, n)
                assert np.array_equal(res[3].view(np.uint16), ref[1].view(np.uint16)), (name, lab.ndim, n)
                _check_direction(res[4], ref[2], lab if lab.ndim == 3 else np.repeat(lab[:, :, None], 3, axis=2), n, name)


@pytest.mark.parametrize("seed,H,W,n", [(71, 128, 160, 14), (72, 250, 200, 60)])
def test_config3_chain_16_directions(kernel_api, seed, H, W, n):
    """BASELINE configs[3]: 16-direction target generation, the direction-difference map of the produced class map
    (17 classes: the '8 of 16 channels' quirk of getDirectionDiffMap.py:69-90) and the 4-connected labelling of the
    interior mask -- device-resident from the label image to the three results"""
    import torch
    from scipy import ndimage as ndi
    from oracle import restate as O
    from cdnet_b200 import synth
    lab = synth.as_uint8_label(synth.instance_map(seed, H, W, n))
    tern, point, direction = O.label_encoding(lab, num_classes=16, literal=False)
    ids = to_dev(kernel_api, torch.from_numpy(lab[:, :, 0].copy())[None])
    g_tern, g_point, g_dir = kernel_api.encode_targets_cuda(ids, True, 16)
    assert np.array_equal(g_tern[0].cpu().numpy(), tern)
    assert np.array_equal(g_point[0].cpu().numpy().view(np.uint16), point.view(np.uint16))
    _check_direction(g_dir[0].cpu().numpy(), direction, lab, 16, "config3 seed %d" % seed)
    # DDM of the class map the device produced (uint8 on the device, like test_dam.py:459 hands it over)
    ddm = kernel_api.ddm_cuda(g_dir.to(torch.uint8), 17)
    ref_ddm = O.generate_dd_map(g_dir[0].cpu().numpy().astype(np.uint8), 17)
    assert np.array_equal(ddm[0].cpu().numpy(), ref_ddm, equal_nan=True)
    interior = (g_tern == 127)
    got, cnt = kernel_api.label_cuda(interior, connectivity=4, return_num=True)
    ref_lab, ref_n = ndi.label(tern == 127)
    assert int(cnt[0]) == ref_n and np.array_equal(got[0].cpu().numpy(), ref_lab)


def test_primitives_fuzz(kernel_api):
    """random small shapes and densities (1 x 1 up to 70 x 90, every W % 4 residue): labelling, hole filling,
    small-object removal, EDT and label dilation against scipy / the oracle"""
    from scipy import ndimage as ndi
    from oracle import restate as O
    rng = np.random.default_rng(2024)
    for it in range(48):
        H, W = int(rng.integers(1, 71)), int(rng.integers(1, 91))
        p = float(rng.choice([0.05, 0.3, 0.5, 0.6, 0.75, 0.95]))
        m = rng.random((H, W)) < p
        if it % 6 == 0:  # blobs instead of salt and pepper
            m = ndi.binary_dilation(rng.random((H, W)) < 0.03, iterations=int(rng.integers(1, 4)))
        tag = (it, H, W, p)
        assert np.array_equal(kernel_api.label(m, connectivity=1), ndi.label(m)[0]), tag
        assert np.array_equal(kernel_api.label(m), O.label8(m)), tag
        assert np.array_equal(kernel_api.binary_fill_holes(m), ndi.binary_fill_holes(m)), tag
        k = int(rng.integers(1, 12))
        assert np.array_equal(kernel_api.remove_small_objects(m, k), O.remove_small_objects(m, k)), tag
        if not m.all():
            assert np.array_equal(kernel_api.distance_transform_edt(m), ndi.distance_transform_edt(m)), tag
        lab = ndi.label(m)[0]
        r = int(rng.integers(1, 3))
        assert np.array_equal(kernel_api.dilation(lab, radius=r), O.dilate(lab, O.disk(r))), tag
        assert np.array_equal(kernel_api.remove_small_objects(lab, k), O.remove_small_objects(lab, k)), tag


def test_ddm_fuzz(kernel_api):
    """random class maps (5 / 9 / 17 classes, ids beyond the table, constant maps) against the restatement"""
    from oracle import restate as O
    rng = np.random.default_rng(77)
    for it in range(36):
        cls = (5, 9, 17)[it % 3]
        H, W = int(rng.integers(1, 60)), int(rng.integers(1, 75))
        x = rng.integers(0, cls + (2 if it % 5 == 0 else 0), size=(H, W)).astype(np.uint8)
        x[rng.random((H, W)) < float(rng.choice([0.0, 0.3, 0.8]))] = 0
        if it % 11 == 0:
            x[:] = x.flat[0]
        got = kernel_api.generate_dd_map(x, cls)
        assert np.array_equal(got, O.generate_dd_map(x, cls), equal_nan=True), (it, cls, H, W)


def test_ddm_bitsliced_shapes(kernel_api):
    """the bit-sliced kernel's shapes (W % 8 == 0): single chunk, exactly / just over 1024 columns (halo lanes),
    several chunks, odd row counts, ids beyond the table (>= n and >= 16), both ring tables"""
    from oracle import restate as O
    rng = np.random.default_rng(5)
    for (H, W, n) in [(37, 64, 9), (16, 1024, 9), (33, 1032, 9), (5, 2000, 9), (40, 96, 5), (18, 1992, 5), (1, 8, 9),
                      (2, 8, 5), (19, 1000, 9), (35, 1016, 5)]:
        lab = rng.integers(0, n, size=(H, W)).astype(np.uint8)
        lab[:, :W // 3] = rng.integers(0, n)
        lab[rng.random((H, W)) < 0.01] = rng.integers(n, 256)
        lab[rng.random((H, W)) < 0.005] = 15
        got = kernel_api.generate_dd_map(lab, n)
        assert np.array_equal(got, O.generate_dd_map(lab, n), equal_nan=True), (H, W, n)


def test_dam_bitsliced_eight_maps(kernel_api):
    """the 8-map form of the bit-sliced kernel on noisy direction maps (every code value in every map), widths
    around the one-chunk limit"""
    from oracle import restate as O
    from cdnet_b200 import synth
    rng = np.random.default_rng(6)
    for (H, W) in [(21, 1000), (34, 1024), (9, 1048)]:
        d = synth.postproc_inputs(300 + H, H, W, 12)
        dcm = d["dcm"].copy()
        noise = rng.random(dcm.shape) < 0.2
        dcm[noise] = rng.integers(0, 12, size=int(noise.sum())).astype(np.uint8)
        pr, pg = d["prob"].copy(), d["prob"].copy()
        ref = O.dam_postprocess(pr, d["point"], dcm, 9, 20, 2, 0, literal=False)
        got = kernel_api.dam_postprocess(pg, d["point"], dcm, 9, 20, 2, 0)
        assert np.array_equal(got, ref["pred_labeled"]), (H, W)
        assert np.array_equal(pg[2], pr[2]), (H, W)  # the boosted boundary channel, written in place by both


def test_run_based_tail_fuzz(kernel_api):
    """fill holes -> remove small -> 8-connected labels -> dilation on adversarial masks (noise of several densities:
    one-pixel runs, nested holes, diagonal-only contacts; blobs with pinholes), ragged shapes (W below / not a multiple
    of 32, one row, one column, wider than one 1024-column chunk), every radius and min_area incl. 1"""
    from scipy import ndimage as ndi
    from oracle import restate as O
    rng = np.random.default_rng(2024)
    shapes = [(1, 1), (1, 37), (40, 1), (7, 31), (9, 32), (23, 33), (50, 64), (31, 100), (64, 257), (12, 1024), (6, 1100),
              (5, 2100), (3, 3000)]
    for it, (H, W) in enumerate(shapes * 2):
        kind = it % 4
        if kind == 0:
            m = rng.random((H, W)) < float(rng.choice([0.3, 0.5, 0.7]))
        elif kind == 1:
            m = ndi.binary_dilation(rng.random((H, W)) < 0.03, iterations=int(rng.integers(1, 5)))
            m &= rng.random((H, W)) < 0.95
        elif kind == 2:
            m = (np.add.outer(np.arange(H), np.arange(W)) % 2 == 0)  # checkerboard: diagonal contacts only
            m &= rng.random((H, W)) < 0.9
        else:
            m = np.ones((H, W), bool)
            m[rng.random((H, W)) < 0.1] = False
        prob = np.zeros((3, H, W), np.float32)
        prob[0] = 0.5
        prob[1] = m
        for radius in (0, 1, 2):
            min_area = int(rng.choice([1, 2, 5, 20]))
            ref = O.plain_postprocess(prob.copy(), min_area, radius, 0)
            got = kernel_api.plain_postprocess(prob.copy(), min_area, radius, 0)
            ref = ref["pred_labeled"] if isinstance(ref, dict) else ref
            assert got.dtype == ref.dtype and np.array_equal(got, ref), (it, H, W, kind, radius, min_area,
                                                                        int((got != ref).sum()))


def test_run_based_tail_frame_cases(kernel_api):
    """background and the image frame (the run-based tail hangs frame-touching background below an 'outside' root):
    all background, all foreground, a ring whose inside reaches the frame only through a corner pixel / a diagonal gap
    (background is 4-connected for binary_fill_holes: a diagonal gap does not open a hole), holes on the first and
    last row / column, every radius"""
    from oracle import restate as O
    cases = {}
    for (H, W) in ((40, 70), (33, 1030)):
        z = np.zeros((H, W), bool)
        cases["zeros_%d" % W] = z
        cases["ones_%d" % W] = ~z
        ring = z.copy()
        ring[5:25, 5:30] = True
        ring[8:22, 8:27] = False
        cases["ring_%d" % W] = ring
        gap = ring.copy()
        gap[5, 5] = False                      # corner of the ring removed: the hole touches outside diagonally only
        cases["ring_diag_gap_%d" % W] = gap
        leak = ring.copy()
        leak[5:8, 15] = False                  # a real 4-connected leak: no hole any more
        cases["ring_leak_%d" % W] = leak
        edge = z.copy()
        edge[0:12, 40:60] = True
        edge[0:6, 45:50] = False               # notch open to the first row: not a hole
        edge[8:10, 52:55] = False              # a hole next to it
        edge[H - 10:H, 0:15] = True
        edge[H - 5:H - 2, 0:4] = False         # notch open to the first column
        edge[H - 8:H - 6, 6:9] = False         # hole
        edge[10:30, W - 12:W] = True
        edge[15:20, W - 3:W] = False           # notch open to the last column
        cases["frame_notches_%d" % W] = edge
        corner = ~z
        corner[0, 0] = False                   # one background pixel, in the corner
        corner[H // 2, W // 2] = False         # and one enclosed
        cases["corner_%d" % W] = corner
    for name, m in cases.items():
        prob = np.zeros((3,) + m.shape, np.float32)
        prob[0] = 0.5
        prob[1] = m
        for radius in (0, 1, 2):
            ref = O.plain_postprocess(prob.copy(), 4, radius, 0)
            ref = ref["pred_labeled"] if isinstance(ref, dict) else ref
            got = kernel_api.plain_postprocess(prob.copy(), 4, radius, 0)
            assert np.array_equal(got, ref), (name, radius, int((got != ref).sum()))


@pytest.mark.parametrize("B,H,W", [(5, 1000, 96), (12, 800, 64)])
def test_run_based_tail_block_rows(kernel_api, B, H, W):
    """batches with 4 736+ rows switch k_rle_pack_link to 16-row blocks (shared-memory union-find over more rows, fewer
    seams for k_rle_link; CDNET_RLE_PACK_ROWS=32 for 32-row blocks); labels must not change"""
    import torch
    from scipy import ndimage as ndi
    from oracle import restate as O
    rng = np.random.default_rng(B * 1000 + W)
    prob = np.zeros((B, 3, H, W), np.float32)
    prob[:, 0] = 0.5
    for b in range(B):
        if b % 3 == 0:
            m = rng.random((H, W)) < 0.55                      # one-pixel structure, nested holes
        elif b % 3 == 1:
            m = ndi.binary_dilation(rng.random((H, W)) < 0.02, iterations=3) & (rng.random((H, W)) < 0.97)
        else:
            yy, xx = np.mgrid[0:H, 0:W]                        # a snake through every block seam + pinholes
            m = (((yy // 5) % 2 == 0) | ((xx < 3) & ((yy // 10) % 2 == 0)) | ((xx >= W - 3) & ((yy // 10) % 2 == 1)))
            m &= rng.random((H, W)) < 0.98
        prob[b, 1] = m
    got, _ = kernel_api.plain_postprocess_cuda(to_dev(kernel_api, torch.from_numpy(prob)), 5, 2, 0)
    got = got.cpu().numpy()
    for b in range(B):
        ref = O.plain_postprocess(prob[b].copy(), 5, 2, 0)
        ref = ref["pred_labeled"] if isinstance(ref, dict) else ref
        assert np.array_equal(got[b], ref), (b, int((got[b] != ref).sum()))


def test_edt_large_components(kernel_api):
    """the distance transform away from nucleus scale: a 1000-pixel-wide blob, a tile with a single background pixel,
    columns / rows without any, a frame-to-frame band -- the scans beyond the fast kernels' cut-off (csrc/edt.cu)"""
    from scipy import ndimage as ndi
    yy, xx = np.mgrid[0:260, 0:1200]
    cases = {
        "blob_1000_wide": ((xx - 600) ** 2 / 500.0 ** 2 + (yy - 130) ** 2 / 120.0 ** 2) <= 1.0,
        "one_zero": np.ones((150, 210), bool),
        "band": np.zeros((300, 180), bool),
        "columns_without_zero": np.ones((90, 140), bool),
    }
    cases["one_zero"][17, 33] = False
    cases["band"][:, 20:160] = True
    cases["columns_without_zero"][40, ::7] = False
    for name, m in cases.items():
        got = kernel_api.distance_transform_edt(m)
        ref = ndi.distance_transform_edt(m)
        assert got.dtype == np.float64 and np.array_equal(got, ref), (name, float(np.abs(got - ref).max()))


@pytest.mark.simt_skip
def test_process_large_blob(kernel_api):
    """postproc_other.process on a tile that holds a 1000-pixel-wide component next to ordinary nuclei"""
    from oracle import restate as O
    from cdnet_b200 import synth
    H, W = 700, 1300
    yy, xx = np.mgrid[0:H, 0:W]
    m = ((xx - 640) ** 2 / 520.0 ** 2 + (yy - 330) ** 2 / 260.0 ** 2) <= 1.0
    small = synth.instance_map(5, H, W, 150) > 0
    m |= small
    m[300:310, 600:700] = False  # a hole in the blob
    a = (m.astype(np.uint8) * 255)
    ref = O.process(a.copy(), "modelName", min_size=10, literal=False)
    got = kernel_api.process(a.copy(), "modelName", min_size=10)
    assert np.array_equal(got, ref), int((got != ref).sum())


def test_process_fuzz(kernel_api):
    """postproc_other.process on random masks: watershed branch and the no-watershed head, several min_size"""
    from scipy import ndimage as ndi
    from oracle import restate as O
    rng = np.random.default_rng(99)
    for it in range(24):
        H, W = int(rng.integers(6, 64)), int(rng.integers(6, 80))
        seeds = rng.random((H, W)) < 0.02
        m = ndi.binary_dilation(seeds, iterations=int(rng.integers(2, 6)))
        m &= rng.random((H, W)) < 0.97  # pinholes
        if m.all() or not m.any():
            continue
        ms = int(rng.choice([1, 5, 10]))
        src = m.astype(np.uint8) * 255
        for mode in ("modelName", "unet"):
            ref = O.process(src.copy(), mode, min_size=ms, literal=False)
            got = kernel_api.process(src.copy(), mode, min_size=ms)
            assert got.dtype == ref.dtype and np.array_equal(got, ref), (it, H, W, ms, mode)


def test_label_encoding_fuzz(kernel_api):
    """target transform on random instance maps of ragged sizes (8 and 16 direction classes)"""
    from oracle import restate as O
    from cdnet_b200 import synth
    rng = np.random.default_rng(4242)
    for it in range(10):
        H, W = int(rng.integers(24, 100)), int(rng.integers(24, 120))
        n = int(rng.integers(1, max(2, H * W // 350)))
        classes = 8 if it % 2 == 0 else 16
        lab = synth.as_uint8_label(synth.instance_map(9000 + it, H, W, n, axes=(3, 9)))
        ref = O.label_encoding(lab, num_classes=classes, literal=False)
        res = kernel_api.LabelEncoding(3, 1, 1, num_classes=classes)((None, None, lab.copy()))
        assert np.array_equal(np.asarray(res[2]), ref[0]), (it, H, W)
        assert np.array_equal(res[3].view(np.uint16), ref[1].view(np.uint16)), (it, H, W)
        _check_direction(res[4], ref[2], lab, classes, "fuzz %d" % it)


def test_label_multi_valued_image(kernel_api):
    """skimage.measure.label on an id image: 8-connected components of equal value, raster-first numbering"""
    from oracle import restate as O
    from cdnet_b200 import synth
    rng = np.random.default_rng(31)
    imgs = [ids for _, ids in synth.label_edge_cases()]
    imgs += [synth.as_uint8_label(synth.instance_map(800 + i, 90 + 7 * i, 101 + 3 * i, 20))[:, :, 0] for i in range(3)]
    imgs += [rng.integers(0, 4, size=(33, 47)).astype(np.uint8), rng.integers(0, 3, size=(1, 29)).astype(np.uint8),
             rng.integers(0, 3, size=(31, 1)).astype(np.uint8), np.full((5, 6), 7, np.uint8)]
    for i, ids in enumerate(imgs):
        got, n = kernel_api.label(ids, return_num=True)
        ref = O.label8_values(ids) if np.unique(ids[ids != 0]).size > 1 else O.label8(ids)
        assert got.dtype == np.int64 and np.array_equal(got, ref), i
        assert n == int(ref.max())


def test_label_encoding_out_c_1(kernel_api):
    """my_transforms_direction.LabelEncoding with out_c != 3 (options.py:42, multi_class off; :721-739): no boundary
    class, undilated instances; the restatement is pinned to the verbatim reference for the same inputs in
    tests/test_oracle_vs_reference.py"""
    from oracle import restate as O
    from cdnet_b200 import synth
    lab = synth.as_uint8_label(synth.instance_map(779, 70, 90, 6))
    binary = np.repeat(((lab[:, :, 0] > 0) * 255).astype(np.uint8)[:, :, None], 3, axis=2)
    shifted = binary.copy()
    shifted[:, :, 1] = np.roll(binary[:, :, 0], 5, axis=1)   # channel 1 matters for a {0,255} label (:730-731)
    cases = [("instance", lab, 8), ("instance16", lab, 16), ("binary", binary, 8), ("binary_ch1", shifted, 8),
             ("dense", synth.as_uint8_label(synth.instance_map(23, 250, 300, 330)), 8)]
    cases += [(name, np.repeat(ids[:, :, None], 3, axis=2), 8) for name, ids in synth.label_edge_cases()]
    for name, img, n in cases:
        for dd in (1, 0):
            ref = O.label_encoding(img.copy(), out_c=1, do_direction=dd, num_classes=n, literal=False)
            res = kernel_api.LabelEncoding(1, 1, dd, num_classes=n)((None, None, img.copy()))
            assert len(res) == (5 if dd else 3), name
            assert np.array_equal(np.asarray(res[2]), ref[0]), name
            if dd:
                assert np.array_equal(res[3].view(np.uint16), ref[1].view(np.uint16)), name
                _check_direction(res[4], ref[2], img, n, name, out_c=1)
    with pytest.raises(IndexError):
        kernel_api.LabelEncoding(1, 1, 1)((None, None, lab[:, :, 0].copy()))


@pytest.mark.parametrize("seed,H,W,n", [(781, 130, 150, 16), (782, 64, 200, 10)])
def test_dam_postprocess_voting_first(kernel_api, seed, H, W, n):
    """`voting_firt = 1` (test_dam.py:471-477): TTA direction voting on the device, then the single-map pipeline"""
    from oracle import restate as O
    from cdnet_b200 import synth
    d = synth.postproc_inputs(seed, H, W, n)
    for pp in (0, 1):
        p_ref, p_got = d["prob"].copy(), d["prob"].copy()
        ref = O.dam_postprocess(p_ref, d["point"], d["dcm"], 9, 20, 2, pp, literal=False, voting_first=True)["pred_labeled"]
        got = kernel_api.dam_postprocess(p_got, d["point"], d["dcm"], 9, 20, 2, pp, voting_first=True)
        assert got.dtype == ref.dtype and np.array_equal(got, ref), (seed, pp)
        assert np.array_equal(p_ref.view(np.uint32), p_got.view(np.uint32))  # prob_maps[2] updated in place (:536)


def test_postprocess_with_unet_model_mode(kernel_api):
    """postproc = 1 with model_mode 'unet': process() skips the watershed (postproc_other.py:35,50-54)"""
    from oracle import restate as O
    from cdnet_b200 import synth
    d = synth.postproc_inputs(783, 110, 140, 14)
    ref = O.plain_postprocess(d["prob"].copy(), 20, 2, 1, model_name="unet", literal=False)["pred_labeled"]
    got = kernel_api.plain_postprocess(d["prob"].copy(), 20, 2, 1, model_name="unet")
    assert got.dtype == ref.dtype and np.array_equal(got, ref)
    ref = O.dam_postprocess(d["prob"].copy(), d["point"], d["dcm"], 9, 20, 2, 1, model_name="unet", literal=False)["pred_labeled"]
    got = kernel_api.dam_postprocess(d["prob"].copy(), d["point"], d["dcm"], 9, 20, 2, 1, model_name="unet")
    assert got.dtype == ref.dtype and np.array_equal(got, ref)
    with pytest.raises(NotImplementedError):
        kernel_api.plain_postprocess(d["prob"].copy(), 20, 2, 1, model_name="dcan")


def _widening_golden():
    import torch
    from cdnet_b200 import synth
    z, meta = load_golden("widening")
    t = meta["tta"]
    g = torch.Generator().manual_seed(t["seed"])
    shapes = [(t["H"], t["W"])] * 4 + [(t["W"], t["H"])] * 4
    ml = [torch.randn((3,) + s, generator=g) * 3 for s in shapes]
    pt = [torch.randn((1,) + s, generator=g) for s in shapes]
    dl = [torch.randn((t["C"],) + s, generator=g) * 3 for s in shapes]
    assert synth.digest(*[x.numpy() for x in ml + pt + dl]) == t["digest"], "torch.Generator stream differs from the goldens'"
    return z, meta, ml, pt, dl


def test_widening_golden(kernel_api):
    """the widening features straight against vectors generated from the verbatim reference
    (oracle/make_goldens.py gold_widening): LabelEncoding out_c != 3, voting first, 'unet' mode"""
    from cdnet_b200 import synth
    z, meta, ml, pt, dl = _widening_golden()
    c = meta["c1"]
    lab = synth.as_uint8_label(synth.instance_map(c["seed"], c["H"], c["W"], c["n_target"]))
    binary = np.repeat(((lab[:, :, 0] > 0) * 255).astype(np.uint8)[:, :, None], 3, axis=2)
    binary[:, :, 1] = np.roll(binary[:, :, 0], 5, axis=1)
    assert synth.digest(lab, binary) == c["digest"]
    for name, img in (("inst", lab), ("bin", binary)):
        r = kernel_api.LabelEncoding(1, 1, 1, num_classes=8)((None, None, img.copy()))
        assert np.array_equal(np.asarray(r[2]), z["c1_%s_tern" % name]), name
        assert np.array_equal(r[3].view(np.uint16), z["c1_%s_point" % name].view(np.uint16)), name
        _check_direction(r[4], z["c1_%s_dir" % name].astype(np.int64), img, 8, "golden c1 " + name, out_c=1)
    p = meta["pp"]
    d = synth.postproc_inputs(p["seed"], p["H"], p["W"], p["n_target"])
    assert synth.digest(d["dcm"], d["prob"], d["point"]) == p["digest"]
    for pp in (0, 1):
        got = kernel_api.dam_postprocess(d["prob"].copy(), d["point"], d["dcm"], 9, 20, 2, pp, voting_first=True)
        assert got.dtype == z["vote_pp%d" % pp].dtype and np.array_equal(got, z["vote_pp%d" % pp]), pp
    assert np.array_equal(kernel_api.plain_postprocess(d["prob"].copy(), 20, 2, 1, model_name="unet"), z["unet_plain"])
    assert np.array_equal(kernel_api.dam_postprocess(d["prob"].copy(), d["point"], d["dcm"], 9, 20, 2, 1, model_name="unet"),
                          z["unet_dam"])



# ---- the fused TTA hand-off (csrc/handoff.cu): scalar and 4-pixel forms ----
def _tta_inputs(seed, B, H, W, C):
    import torch
    g = torch.Generator().manual_seed(seed)
    shapes = [(H, W)] * 4 + [(W, H)] * 4
    ml = [torch.randn((B, 3) + s, generator=g) * 3 for s in shapes]
    pt = [torch.randn((B, 1) + s, generator=g) for s in shapes]
    dl = [torch.randn((B, C) + s, generator=g) * 3 for s in shapes]
    return ml, pt, dl


# H % 4 == W % 4 == 0 takes the 4-pixel (128-bit) kernel, everything else the scalar one
@pytest.mark.parametrize("B,H,W,C", [(1, 21, 34, 9), (2, 64, 96, 9), (1, 33, 31, 17), (1, 1, 1, 5), (1, 70, 5, 9),
                                     (1, 36, 100, 17), (2, 100, 36, 5), (1, 4, 4, 9), (1, 8, 132, 9)])
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
            clear = O.tta_variant_to_original(((top[-1] - top[-2]) > 3e-6)[None], v)[0]
            assert np.array_equal(got[v][clear], rd[v][clear]), (b, v)
            assert clear.mean() > 0.99


@pytest.mark.gpu
def test_tta_merge_vs_torch_on_the_same_gpu(cuda_api):
    """The reference runs its soft-max / argmax ON THE GPU (test_dam.py:984-1013): the right witness for the fused
    hand-off's direction classes is torch.softmax / torch.argmax on the same device.  The mismatch count is recorded."""
    import torch
    from conftest import record_parity
    api = cuda_api
    B, H, W, C = 2, 500, 500, 9
    g = torch.Generator(device="cuda").manual_seed(3)
    ml = [torch.randn((B, 3, H, W), device="cuda", generator=g) * 3 for _ in range(8)]
    pt = [torch.randn((B, 1, H, W), device="cuda", generator=g) for _ in range(8)]
    dl = [torch.randn((B, C, H, W), device="cuda", generator=g) * 3 for _ in range(8)]
    prob, point, dcm = api.tta_merge_cuda(ml, pt, dl)
    mism, total, ties = 0, 0, 0
    psum = torch.zeros_like(prob)
    for v in range(8):
        p = torch.softmax(ml[v], dim=1)
        dp = torch.softmax(dl[v], dim=1)
        dp[:, 0] = dp[:, 0] * p[:, 0]
        cls = torch.argmax(dp, dim=1, keepdim=True)

        def back(t):
            if v & 1:
                t = torch.flip(t, dims=(3,))
            if v & 2:
                t = torch.flip(t, dims=(2,))
            if v & 4:
                t = torch.rot90(t, 3, dims=(2, 3))
            return t
        ref = back(cls)[:, 0]
        top2 = back(torch.topk(dp, 2, dim=1).values)
        near = (top2[:, 0] - top2[:, 1]) < 3e-6
        bad = dcm[:, v].long() != ref
        mism += int(bad.sum())
        ties += int((bad & near).sum())
        total += ref.numel()
        psum += back(p)
    record_parity("tta_merge_vs_torch_same_gpu", class_mismatches=mism, of_which_near_ties=ties, px_x_variants=total,
                  max_abs_prob_diff=float((prob - psum / 8).abs().max()))
    assert mism == 0, "direction classes differ from torch's on the same GPU (%d, %d of them at numerical ties)" % (mism, ties)
    assert torch.allclose(prob, psum / 8, rtol=1e-5, atol=1e-7)


def test_hand_off_without_tta(kernel_api):
    """one variant (tta off): soft-max / argmax only, nothing averaged; dcm has one map (test_dam.py:499-502)"""
    from oracle import restate as O
    ml, pt, dl = _tta_inputs(7, 2, 40, 56, 9)
    prob, point, dcm = kernel_api.tta_merge_cuda([to_dev(kernel_api, ml[0])], [to_dev(kernel_api, pt[0])],
                                                 [to_dev(kernel_api, dl[0])])
    assert tuple(dcm.shape) == (2, 1, 40, 56)
    for b in range(2):
        p, q, c = O.variant_probmaps(ml[0][b].numpy(), pt[0][b].numpy(), dl[0][b].numpy())
        assert np.allclose(prob[b].cpu().numpy(), p, rtol=1e-5, atol=1e-7)
        assert np.array_equal(point[b].cpu().numpy(), q)
        z = dl[0][b].numpy().astype(np.float64)
        qq = np.exp(z - z.max(axis=0)) / np.exp(z - z.max(axis=0)).sum(axis=0)
        qq[0] *= p[0]
        top = np.sort(qq, axis=0)
        clear = (top[-1] - top[-2]) > 3e-6
        assert np.array_equal(dcm[b, 0].cpu().numpy().astype(np.int64)[clear], c[0][clear]) and clear.mean() > 0.99


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


def test_tta_merge_golden(kernel_api):
    """the fused TTA hand-off against the vectors of the reference's TTA block executed verbatim"""
    z, meta, ml, pt, dl = _widening_golden()
    prob, point, dcm = kernel_api.tta_merge_cuda([to_dev(kernel_api, t[None]) for t in ml],
                                                 [to_dev(kernel_api, t[None]) for t in pt],
                                                 [to_dev(kernel_api, t[None]) for t in dl])
    assert np.allclose(prob[0].cpu().numpy(), z["tta_prob"], rtol=1e-5, atol=1e-7)
    assert np.array_equal(point[0].cpu().numpy().view(np.uint32), z["tta_point"].view(np.uint32))
    assert (dcm[0].cpu().numpy() != z["tta_dcm"]).mean() < 2e-3  # float near-ties only


@pytest.mark.gpu
def test_device_guard_other_gpu(cuda_api):
    """tensors on cuda:1 while cuda:0 is the current device: the call runs on the tensors' device (api._on_tensor_device)
    and gives the same labels as on cuda:0.  Needs two GPUs (skipped on the single-GPU test box)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from cdnet_b200 import synth
    api = cuda_api
    d = synth.postproc_inputs(5, 160, 200, 20)
    torch.cuda.set_device(0)
    on0 = [torch.from_numpy(np.ascontiguousarray(d[k])).to("cuda:0")[None] for k in ("dcm", "prob", "point")]
    on1 = [t.to("cuda:1") for t in on0]
    lab0, _ = api.dam_postprocess_cuda(*on0, 9, 20, 2, 0)
    assert torch.cuda.current_device() == 0
    lab1, _ = api.dam_postprocess_cuda(*on1, 9, 20, 2, 0)
    assert lab1.device.index == 1 and torch.cuda.current_device() == 0
    assert torch.equal(lab0.cpu(), lab1.cpu())
    plan = api.DamPostprocessPlan(1, 160, 200, 9, 20, 2, 0, device=1)
    plan.h_dcm[0], plan.h_prob[0], plan.h_point[0] = d["dcm"], d["prob"], d["point"]
    assert np.array_equal(plan.run()[0], lab0[0].cpu().numpy())
