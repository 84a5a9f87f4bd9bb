"""Generate tests/golden/*.npz by running the REFERENCE VERBATIM (oracle/ref_loader.py).

TEST INFRASTRUCTURE ONLY; runs only where /root/reference is mounted (the build container):

    python -m oracle.make_goldens            # everything (the 1000x1000 target tile takes minutes)
    python -m oracle.make_goldens --quick    # skip the 1000x1000 target-transform tile

Inputs are NOT stored: they are regenerated from the seed by cdnet_b200/synth.py (bit-exact by
construction) and each file records the sha1 digest of its inputs so a test can verify that.
The 16-direction goldens are produced in a child process with dt_num_classes=16 in the
environment, because the reference freezes that setting at import
(data_prepare/SegFix_offset_helper.py:37-39).
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(REPO, "tests", "golden")
sys.path.insert(0, REPO)

from cdnet_b200 import synth  # noqa: E402
from oracle import ref_loader  # noqa: E402

# (name, seed, H, W, n_target) -- post-processing tiles
P_TILES = [
    ("p_96x128", 11, 96, 128, 9),
    ("p_200x150", 12, 200, 150, 22),
    ("p_256", 100, 256, 256, None),
    ("p_333x517", 13, 333, 517, None),
    ("p_1000", 100, 1000, 1000, None),
]
# target-transform tiles: (name, seed, H, W, n_target)
T_TILES = [
    ("t_64_single", 21, 64, 64, 1),
    ("t_128", 22, 128, 128, 12),
    ("t_256", 5, 256, 256, 45),
    ("t_250x300_dense", 23, 250, 300, 330),  # > 255 nuclei: uint8 id wrap
    ("t_500", 1000, 500, 500, 120),
]
T_BIG = ("t_1000", 0, 1000, 1000, 700)


def _save(name, meta, **arrays):
    os.makedirs(GOLD, exist_ok=True)
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, meta=np.asarray(json.dumps(meta)), **arrays)
    print("wrote %s (%.1f KB)" % (path, os.path.getsize(path) / 1024.0), flush=True)


def gold_ddm(ns):
    rng = np.random.default_rng(42)
    out, meta = {}, {"cases": []}
    for cls in (5, 9, 17):
        for i, (H, W) in enumerate(((37, 53), (64, 64), (1, 9), (9, 1))):
            x = rng.integers(0, cls, size=(H, W)).astype(np.uint8)
            x[rng.integers(0, 3, size=(H, W)) == 0] = 0
            key = "c%d_%d" % (cls, i)
            out[key + "_in"] = x
            out[key + "_out"] = ns.generate_dd_map(x, cls)
            meta["cases"].append([key, cls])
        for key, x in (("c%d_zero" % cls, np.zeros((6, 7), np.uint8)),
                       ("c%d_const" % cls, np.full((6, 7), 1, np.uint8))):
            out[key + "_in"] = x
            out[key + "_out"] = ns.generate_dd_map(x, cls)
            meta["cases"].append([key, cls])
    # a real-looking one
    d = synth.postproc_inputs(100, 256, 256)
    out["tta0_in"] = d["dcm"][0]
    out["tta0_out"] = ns.generate_dd_map(d["dcm"][0], 9)
    meta["cases"].append(["tta0", 9])
    m = rng.integers(-3, 4, size=(2, 6, 8))
    shifts = [(1, 1, 1), (1, 1, 0), (2, 1, 1), (3, 0, 1), (4, 0, 1), (3, 1, 1), (3, 1, 0), (4, 1, 1),
              (1, 2, 3), (4, 0, 0)]
    out["cs_in"] = m
    for i, (dd, s1, s2) in enumerate(shifts):
        out["cs_%d" % i] = ns.circshift(m, dd, s1, s2)
    meta["circshift"] = shifts
    _save("ddm", meta, **out)


def gold_postproc(ns, quick):
    for name, seed, H, W, n in P_TILES:
        t0 = time.time()
        d = synth.postproc_inputs(seed, H, W, n)
        meta = {"seed": seed, "H": H, "W": W, "n_target": n, "direction_classes": 9, "min_area": 20,
                "radius": 2, "digest": synth.digest(d["dcm"], d["prob"], d["point"])}
        arrays = {}
        for pp in (0, 1):
            r = ns.dam_postprocess(d["prob"].copy(), d["point"], d["dcm"], 9, 20, 2, pp)
            arrays["dam_pp%d_labels" % pp] = r["pred_labeled"].astype(np.int32)
            arrays["dam_pp%d_dtype" % pp] = np.asarray(str(r["pred_labeled"].dtype))
            if pp == 0:
                arrays["dam_inside"] = np.packbits(r["pred_inside"])
                arrays["dam_pred2"] = np.packbits(r["pred2"].astype(bool))
                arrays["dam_ddm_mean16"] = np.round(r["prob_direction_maps"][0] * 16).astype(np.uint8)
            r = ns.plain_postprocess(d["prob"].copy(), 20, 2, pp)
            arrays["plain_pp%d_labels" % pp] = r["pred_labeled"].astype(np.int32)
            arrays["plain_pp%d_dtype" % pp] = np.asarray(str(r["pred_labeled"].dtype))
        # watershed order exposure: pixels whose label depends on heap mechanics
        ref_loader.set_watershed_order("heap")
        rh = ns.dam_postprocess(d["prob"].copy(), d["point"], d["dcm"], 9, 20, 2, 1)
        ref_loader.set_watershed_order("stable")
        meta["ws_heap_vs_stable_diff_px"] = int((rh["pred_labeled"] != arrays["dam_pp1_labels"]).sum())
        meta["seconds"] = round(time.time() - t0, 2)
        _save(name, meta, **arrays)


def gold_process(ns):
    """postproc_other.process on raw masks incl. edge cases."""
    out, meta = {}, {"cases": []}
    rng = np.random.default_rng(7)
    cases = {}
    ids = synth.instance_map(31, 180, 220, 40)
    cases["blobs"] = (ids > 0)
    m = np.zeros((60, 80), bool)
    cases["empty"] = m.copy()
    m2 = m.copy(); m2[20:45, 30:60] = True; m2[30:34, 40:44] = False
    cases["holed_rect"] = m2
    m3 = m.copy(); m3[0:25, 0:30] = True; m3[40:60, 60:80] = True
    cases["border"] = m3
    m4 = m.copy(); m4[10:30, 10:30] = True; m4[30:50, 30:50] = True  # diagonal-only contact
    cases["diag"] = m4
    m5 = m.copy(); m5[5:8, 5:8] = True; m5[20, 20] = True
    cases["tiny"] = m5
    yy, xx = np.mgrid[0:90, 0:140]
    dumb = ((yy - 45) ** 2 + (xx - 45) ** 2 <= 28 ** 2) | ((yy - 45) ** 2 + (xx - 95) ** 2 <= 28 ** 2)
    cases["dumbbell"] = dumb
    cases["noise"] = rng.integers(0, 4, size=(70, 90)) > 0
    for key, mask in cases.items():
        for ms in (5, 10):
            src = mask.astype(np.uint8) * 255
            res = ns.process(src.copy(), "modelName", min_size=ms)
            out["%s_ms%d" % (key, ms)] = res.astype(np.int32)
        res = ns.process(mask.astype(np.uint8) * 255, "unet", min_size=10)
        out["%s_unet" % key] = res.astype(np.int32)
        out["%s_in" % key] = np.packbits(mask)
        meta["cases"].append([key, list(mask.shape)])
    _save("process", meta, **out)


def gold_centre(ns):
    out, meta = {}, {"cases": []}
    ids = synth.instance_map(33, 120, 160, 14)
    for k in range(1, int(ids.max()) + 1):
        m = (ids == k).astype(np.int64)
        c = ns.get_centerpoint2(m, m.shape[0], m.shape[1])
        out["c_%d" % k] = np.asarray(c, dtype=np.int64)
    out["ids"] = ids
    meta["n"] = int(ids.max())
    _save("centre", meta, **out)


def _t_case(ns, name, seed, H, W, n, num_classes):
    t0 = time.time()
    ids = synth.instance_map(seed, H, W, n)
    lab = synth.as_uint8_label(ids)
    res = ns.LabelEncoding(3, 1, 1)((None, None, lab))
    meta = {"seed": seed, "H": H, "W": W, "n_target": n, "num_classes": num_classes,
            "n_instances": int(ids.max()), "digest": synth.digest(lab),
            "seconds": round(time.time() - t0, 2)}
    assert res[4].dtype == np.int64 and res[3].dtype == np.float16
    _save(name, meta, ternary=np.asarray(res[2]), point=res[3], direction=res[4].astype(np.uint8))


def gold_targets(ns, quick, num_classes):
    sfx = "" if num_classes == 8 else "_d%d" % num_classes
    tiles = T_TILES if num_classes == 8 else T_TILES[1:3]
    for name, seed, H, W, n in tiles:
        _t_case(ns, name + sfx, seed, H, W, n, num_classes)
    if num_classes == 8:
        # three-class input (values {0,255}) -> measure.label branch, my_transforms_direction.py:763-774
        ids = synth.instance_map(24, 128, 160, 14)
        lab3 = np.repeat(((ids > 0) * 255).astype(np.uint8)[:, :, None], 3, axis=2)
        res = ns.LabelEncoding(3, 1, 1)((None, None, lab3))
        _save("t_128x160_threeclass", {"seed": 24, "H": 128, "W": 160, "n_target": 14,
                                       "num_classes": 8, "digest": synth.digest(lab3)},
              ternary=np.asarray(res[2]), point=res[3], direction=res[4].astype(np.uint8))
        if not quick:
            _t_case(ns, T_BIG[0], *T_BIG[1:], num_classes=8)


# instance metrics: (name, seed, H, W, n_target, mode)
M_CASES = [
    ("m_64x80_unrelated", 1, 64, 80, 6, 0),
    ("m_128", 2, 128, 128, 20, 1),
    ("m_200x150", 3, 200, 150, 40, 1),
    ("m_256_unrelated", 4, 256, 256, 60, 0),
    ("m_500", 6, 500, 500, 150, 1),
    ("m_1000", 7, 1000, 1000, 700, 1),
]


def gold_metrics():
    """stats_utils.py executed verbatim (it only needs numpy + scipy) on synth.metric_pair inputs."""
    import contextlib
    import importlib.util
    import io
    spec = importlib.util.spec_from_file_location("ref_stats_utils", os.path.join(ref_loader.find_reference(), "stats_utils.py"))
    R = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(R)
    out, meta = {}, {"cases": []}
    for name, seed, H, W, n, mode in M_CASES:
        t0 = time.time()
        true, pred = synth.metric_pair(seed, H, W, n, mode)
        with contextlib.redirect_stdout(io.StringIO()):
            aji = R.get_fast_aji(true, pred)
        out[name + "_aji"] = np.asarray(aji, dtype=np.float64)
        out[name + "_aji_plus"] = np.asarray(R.get_fast_aji_plus(true, pred), dtype=np.float64)
        for mi in (0.5, 0.3):
            (dq, sq, pq), (pt, pp, ut, up) = R.get_fast_pq(true, pred, mi)
            tag = name + "_pq%02d" % int(mi * 10)
            out[tag] = np.asarray([dq, sq, pq], dtype=np.float64)
            out[tag + "_paired_true"] = np.asarray(pt, dtype=np.int64)
            out[tag + "_paired_pred"] = np.asarray(pp, dtype=np.int64)
            out[tag + "_unpaired_true"] = np.asarray(ut, dtype=np.int64)
            out[tag + "_unpaired_pred"] = np.asarray(up, dtype=np.int64)
        out[name + "_dice1"] = np.asarray(R.get_dice_1(true, pred), dtype=np.float64)
        out[name + "_dice2"] = np.asarray(R.get_fast_dice_2(true, pred), dtype=np.float64)
        raw = (true.astype(np.int64) * 3 + (true > 0) * 5).astype(np.int32)   # non-contiguous ids
        out[name + "_remap"] = R.remap_label(raw).astype(np.int32)
        out[name + "_remap_by_size"] = R.remap_label(raw, by_size=True).astype(np.int32)
        meta["cases"].append({"name": name, "seed": seed, "H": H, "W": W, "n_target": n, "mode": mode,
                              "digest": synth.digest(true, pred)})
        print("%s: %.1fs (aji %.4f)" % (name, time.time() - t0, float(aji[0])), flush=True)
    _save("metrics", meta, **out)


def gold_training(ns):
    """Quantiser operators (DTOffsetHelper), the direction one-hot block (train_util_dam.py:123-142) and
    my_transforms.LabelEncoding without direction, all executed verbatim (8 direction classes)."""
    import torch
    H = ns.DTOffsetHelper
    d = synth.training_inputs()
    out, meta = {}, {"digest": synth.digest(*[d[k] for k in sorted(d)]), "plain": []}
    for tag in ("32", "64"):
        a = d["angle" + tag]
        for n in (8, 4):
            s, i = H.align_angle(a.copy(), num_classes=n)
            out["align%d_np%s_s" % (n, tag)], out["align%d_np%s_i" % (n, tag)] = s, i
            s, i = H.align_angle(torch.from_numpy(a.copy()), num_classes=n, return_tensor=True)
            out["align%d_pt%s_s" % (n, tag)], out["align%d_pt%s_i" % (n, tag)] = s.numpy(), i.numpy()
            out["a2v%d_np%s" % (n, tag)] = H.angle_to_vector(a.copy(), num_classes=n)
            out["a2v%d_pt%s" % (n, tag)] = H.angle_to_vector(torch.from_numpy(a.copy()), num_classes=n,
                                                             return_tensor=True).numpy()
        out["v2l8_np" + tag] = H.vector_to_label(d["vec" + tag].copy(), num_classes=8)
        out["v2l8_roundtrip" + tag] = H.vector_to_label(out["a2v8_np" + tag].copy(), num_classes=8)
    for C in (4, 5, 8, 9, 16, 17, 32):
        out["l2v%d" % C] = H.label_to_vector(torch.from_numpy(d["labels17"].copy()), num_classes=C).numpy()
    out["onehot9"] = ns.direction_one_hot(torch.from_numpy(d["onehot_dir"].copy()),
                                          torch.from_numpy(d["onehot_target"].copy()), 9).numpy()
    for name, ids in synth.label_edge_cases():
        lab = np.repeat(ids[:, :, None], 3, axis=2)
        lab[:, :, 1] = np.where(ids > 0, 0, 255)[..., ::-1]  # a second channel that matters for out_c != 3
        for out_c in (3, 1):
            out["plain_%s_c%d" % (name, out_c)] = np.asarray(ns.LabelEncodingPlain(out_c, 1, 0)((None, None, lab.copy()))[2])
        meta["plain"].append(name)
    _save("training", meta, **out)


PLAINDIR_CASES = [("pd_64", 3, 64, 64, 5), ("pd_96x120", 4, 96, 120, 14), ("pd_150x130", 5, 150, 130, 30),
                  ("pd_256_dense", 8, 256, 256, 110)]


def gold_plain_direction(ns):
    """my_transforms.LabelEncoding(3, 1, do_direction=1) executed verbatim (my_transforms.py:661-836; peak_local_max
    from oracle/refshim/skimage/feature.py, 8 direction classes) on instance-id and {0,255} labels."""
    out, meta = {}, {"cases": []}
    for name, seed, H, W, n in PLAINDIR_CASES:
        ids = synth.instance_map(seed, H, W, n)
        labs = {"inst": synth.as_uint8_label(ids), "bin": np.repeat(((ids > 0) * 255).astype(np.uint8)[:, :, None], 3, axis=2)}
        for kind, lab in labs.items():
            r = ns.LabelEncodingPlain(3, 1, 1)((None, None, lab.copy()))
            out["%s_%s_ternary" % (name, kind)] = np.asarray(r[2])
            out["%s_%s_point" % (name, kind)] = r[3]
            out["%s_%s_direction" % (name, kind)] = r[4].astype(np.int8)
        meta["cases"].append({"name": name, "seed": seed, "H": H, "W": W, "n": n, "digest": synth.digest(labs["inst"])})
    _save("plaindir", meta, **out)


def gold_widening(ns):
    """Round-1 widening features executed verbatim: the TTA block + get_probmaps (test_dam.py:314-450, :930-1034),
    LabelEncoding with out_c != 3 (my_transforms_direction.py:721-739), the `voting_firt` switch (:471) and
    model_mode 'unet' (postproc_other.py:35).  Inputs come from seeds (synth / torch.Generator)."""
    import torch
    out, meta = {}, {}
    g = torch.Generator().manual_seed(2025)
    H, W, C = 36, 52, 9
    shapes = [(H, W)] * 4 + [(W, H)] * 4
    ml = [torch.randn((3,) + s, generator=g) * 3 for s in shapes]
    pt = [torch.randn((1,) + s, generator=g) for s in shapes]
    dl = [torch.randn((C,) + s, generator=g) * 3 for s in shapes]
    prob, point, dcm = ns.tta_merge(ml, pt, dl)
    out["tta_prob"], out["tta_point"], out["tta_dcm"] = prob, point, dcm.astype(np.uint8)
    meta["tta"] = {"seed": 2025, "H": H, "W": W, "C": C,
                   "digest": synth.digest(*[t.numpy() for t in ml + pt + dl])}
    lab = synth.as_uint8_label(synth.instance_map(779, 70, 90, 6))
    binary = np.repeat(((lab[:, :, 0] > 0) * 255).astype(np.uint8)[:, :, None], 3, axis=2)
    binary[:, :, 1] = np.roll(binary[:, :, 0], 5, axis=1)
    for name, img in (("inst", lab), ("bin", binary)):
        r = ns.LabelEncoding(1, 1, 1)((None, None, img.copy()))
        out["c1_%s_tern" % name], out["c1_%s_point" % name], out["c1_%s_dir" % name] = np.asarray(r[2]), r[3], r[4]
    meta["c1"] = {"seed": 779, "H": 70, "W": 90, "n_target": 6, "digest": synth.digest(lab, binary)}
    d = synth.postproc_inputs(780, 130, 150, 16)
    meta["pp"] = {"seed": 780, "H": 130, "W": 150, "n_target": 16, "digest": synth.digest(d["dcm"], d["prob"], d["point"])}
    for pp in (0, 1):
        out["vote_pp%d" % pp] = ns.dam_postprocess(d["prob"].copy(), d["point"], d["dcm"], 9, 20, 2, pp,
                                                   voting_first=True)["pred_labeled"]
    out["unet_plain"] = ns.plain_postprocess(d["prob"].copy(), 20, 2, 1, model_name="unet")["pred_labeled"]
    out["unet_dam"] = ns.dam_postprocess(d["prob"].copy(), d["point"], d["dcm"], 9, 20, 2, 1, model_name="unet")["pred_labeled"]
    _save("widening", meta, **out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default="")
    ap.add_argument("--child16", action="store_true")
    a = ap.parse_args()
    if a.only == "metrics":
        gold_metrics()
        return
    if a.child16:
        assert os.environ.get("dt_num_classes") == "16"
        ns = ref_loader.load()
        assert ns.DTOffsetConfig.num_classes == 16
        gold_targets(ns, True, 16)
        return
    ns = ref_loader.load()
    assert ns.DTOffsetConfig.num_classes == 8
    todo = a.only.split(",") if a.only else ["ddm", "process", "centre", "postproc", "targets", "t16", "metrics", "training", "widening", "plaindir"]
    if "ddm" in todo:
        gold_ddm(ns)
    if "process" in todo:
        gold_process(ns)
    if "centre" in todo:
        gold_centre(ns)
    if "postproc" in todo:
        gold_postproc(ns, a.quick)
    if "targets" in todo:
        gold_targets(ns, a.quick, 8)
    if "metrics" in todo:
        gold_metrics()
    if "training" in todo:
        gold_training(ns)
    if "widening" in todo:
        gold_widening(ns)
    if "plaindir" in todo:
        gold_plain_direction(ns)
    if "t16" in todo:
        env = dict(os.environ, dt_num_classes="16")
        subprocess.check_call([sys.executable, "-m", "oracle.make_goldens", "--child16"], env=env,
                              cwd=REPO)


if __name__ == "__main__":
    main()
