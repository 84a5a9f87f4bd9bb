"""Load the reference (honglianghe/CDNet) VERBATIM from /root/reference under import shims.

TEST INFRASTRUCTURE ONLY, and only usable where the reference tree is mounted (the build
container).  It is used to (a) validate `oracle/restate.py` and (b) generate `tests/golden/`
(`oracle/make_goldens.py`).  Nothing here is copied from the reference: modules are imported from
where they lie, and the two inline post-processing blocks (`test_dam.py:455-563`,
`test.py:270-295`) are read from the reference files at run time, dedented and `exec`-ed.

Shims: np.float/np.int aliases (after scipy import), collections.Iterable,
stand-in packages from oracle/refshim (scikit-image subset on scipy, empty stubs).
`dt_num_classes` must be in the environment BEFORE the first call (the reference freezes it at
import, data_prepare/SegFix_offset_helper.py:37-39): run 8-way and 16-way in separate processes.
"""
import collections
import collections.abc
import importlib
import os
import sys
import textwrap
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SHIM_DIR = os.path.join(HERE, "refshim")


def find_reference():
    """the mounted tree (build container) or the unmodified copy oracle/install_ref.py put under baseline/_ref
    (what travels to the GPU box)"""
    for p in (os.environ.get("CDNET_REF"), "/root/reference", os.path.join(os.path.dirname(HERE), "baseline", "_ref")):
        if p and os.path.isfile(os.path.join(p, "postproc_other.py")):
            return p
    return None


def available():
    return find_reference() is not None


_ns = None


def load():
    """Returns a namespace with the reference's hot-path callables."""
    global _ns
    if _ns is not None:
        return _ns
    root = find_reference()
    if root is None:
        raise RuntimeError("reference tree not found (set CDNET_REF or mount /root/reference)")
    import scipy.ndimage  # noqa: F401  (must precede the numpy aliases)
    if not hasattr(np, "float"):
        np.float = float
    if not hasattr(np, "int"):
        np.int = int
    if not hasattr(collections, "Iterable"):
        collections.Iterable = collections.abc.Iterable
    repo = os.path.dirname(HERE)
    for p in (repo, SHIM_DIR, root):
        if p not in sys.path:
            sys.path.insert(0, p)
    ns = types.SimpleNamespace(root=root)
    ddm = importlib.import_module("data_prepare.getDirectionDiffMap")
    helper = importlib.import_module("data_prepare.SegFix_offset_helper")
    pp = importlib.import_module("postproc_other")
    mtd = importlib.import_module("my_transforms_direction")
    ns.generate_dd_map = ddm.generate_dd_map
    ns.circshift = ddm.circshift
    ns.DTOffsetHelper = helper.DTOffsetHelper
    ns.DTOffsetConfig = helper.DTOffsetConfig
    ns.Sobel = helper.Sobel
    ns.label_to_vector_mapping = helper.label_to_vector_mapping
    ns.process = pp.process
    ns.LabelEncoding = mtd.LabelEncoding
    ns.get_centerpoint2 = mtd.get_centerpoint2
    ns.dam_postprocess = lambda *a, **k: _dam_postprocess(ns, *a, **k)
    ns.plain_postprocess = lambda *a, **k: _plain_postprocess(ns, *a, **k)
    ns.DcmVoting2 = _load_function(root, "utils.py", "DcmVoting2")
    ns.LabelEncodingPlain = importlib.import_module("my_transforms").LabelEncoding  # my_transforms.py:661-837
    ns.direction_one_hot = lambda *a, **k: _direction_one_hot(ns, *a, **k)
    ns.tta_merge = lambda *a, **k: _tta_merge(ns, *a, **k)
    _ns = ns
    return ns


def set_watershed_order(order):
    """'stable' (canonical) or 'heap' (recalled scikit-image heap mechanics)."""
    load()
    import skimage.segmentation as seg
    assert order in ("stable", "heap")
    seg._ORDER[0] = order


def _read_block(root, fname, first, last):
    with open(os.path.join(root, fname), "r", encoding="utf-8") as f:
        lines = f.readlines()
    return textwrap.dedent("".join(lines[first - 1:last]))


def _load_function(root, fname, name, extra=None):
    """exec one top-level function of a reference file that cannot be imported as a module."""
    with open(os.path.join(root, fname), "r", encoding="utf-8") as f:
        lines = f.readlines()
    start = next(i for i, l in enumerate(lines) if l.startswith("def %s(" % name))
    end = start + 1
    while end < len(lines) and (lines[end].strip() == "" or lines[end][0] in " \t"):
        end += 1
    g = {"np": np}
    g.update(extra or {})
    exec(compile("".join(lines[start:end]), fname, "exec"), g)
    return g[name]


class _Sink(object):
    """swallows cv2.imwrite / io.imsave / Image.save calls inside the inline blocks"""

    def __getattr__(self, name):
        return lambda *a, **k: None


def _dam_postprocess(ns, prob_maps, point_maps, dcm_tta, direction_classes=9, min_area=20,
                     radius=2, postproc=0, model_name="modelName", voting_first=False):
    """Runs test_dam.py:455-563 verbatim.  dcm_tta: 8 maps [8,H,W] (uint8); prob_maps f32 [3,H,W]
    (modified in place like the reference does, :536); point_maps f32 [1,H,W].
    Returns dict(pred_labeled, pred_inside, pred2, prob_direction_maps)."""
    import skimage.morphology as morph
    from skimage import measure
    from scipy import ndimage as ndi
    opt = types.SimpleNamespace(
        model={"mseloss": 1, "direction": 1, "modelName": model_name},
        direction_classes=direction_classes,
        post={"min_area": min_area, "radius": radius, "postproc": postproc},
        transform={"test": {}})
    names = ["prob_dcm", "prob_dcm_hf", "prob_dcm_vf", "prob_dcm_hvf", "prob_dcm_r90",
             "prob_dcm_r90_hf", "prob_dcm_r90_vf", "prob_dcm_r90_hvf"]
    g = {"np": np, "cv2": _Sink(), "io": _Sink(), "morph": morph, "measure": measure, "ndi": ndi,
         "postproc_other": importlib.import_module("postproc_other"),
         "generate_dd_map": ns.generate_dd_map, "DcmVoting2": ns.DcmVoting2, "opt": opt,
         "seg_folder": "", "name": "tile", "branch": "", "Image": _Sink(),
         "save_view_detail_dir": "", "label_img": None,
         "prob_maps": prob_maps, "point_maps": point_maps, "multiple_number": 1.0,
         "label_img_instance": np.zeros((1, 1)), "count_pred_list": [], "count_label_list": [],
         "print": lambda *a, **k: None}
    for i, n in enumerate(names):
        g[n] = np.asarray(dcm_tta[i])[None]
    code = _read_block(ns.root, "test_dam.py", 455, 563)
    if voting_first:
        # the block's own hard-wired switch (test_dam.py:471): DcmVoting2 first, then ONE direction-difference map
        assert code.count("voting_firt = 0") == 1
        code = code.replace("voting_firt = 0", "voting_firt = 1")
    exec(compile(code, "test_dam.py:455-563", "exec"), g)
    return {"pred_labeled": g["pred_labeled"], "pred_inside": g["pred_inside"],
            "pred2": g["pred2"], "prob_direction_maps": g["prob_direction_maps"]}


class _Tok(object):
    """stands in for the PIL image / input tensor of one TTA variant inside the verbatim TTA block"""

    def __init__(self, key=()):
        self.key = key

    def rotate(self, angle, expand=False):
        assert angle == 90 and expand
        return _Tok(self.key + ("r90",))

    def transpose(self, how):
        return _Tok(self.key + (int(how),))

    def unsqueeze(self, dim):
        return self

    def cuda(self):
        return self


def _tta_merge(ns, mask_logits, point, dir_logits):
    """Runs the reference's test-time augmentation verbatim: `get_probmaps` (test_dam.py:930-1034: softmax,
    direction[0] *= mask[0], argmax) for each of the 8 variants and the un-flip / un-rotate / average block
    (test_dam.py:314-450).  Variant order: id, hf, vf, hvf, r90, r90_hf, r90_vf, r90_hvf; mask_logits[v] torch
    float32 [3,h,w], point[v] [1,h,w], dir_logits[v] [C,h,w] in the variant's own frame (the model's outputs).
    Returns (prob_maps f32 [3,H,W], point_maps f32 [1,H,W], the 8 un-flipped direction-class maps [8,H,W])."""
    import copy
    import torch
    import torch.nn.functional as F
    from PIL import Image
    LR, TB = int(Image.FLIP_LEFT_RIGHT), int(Image.FLIP_TOP_BOTTOM)
    keys = [(), (LR,), (TB,), (LR, TB), ("r90",), ("r90", LR), ("r90", TB), ("r90", LR, TB)]
    index = {k: i for i, k in enumerate(keys)}
    opt = types.SimpleNamespace(model={"modelName": "CDNet", "mseloss": 1, "direction": 1, "multi_class": True},
                                test={"patch_size": 0, "overlap": 0})
    get_probmaps = _load_function(ns.root, "test_dam.py", "get_probmaps",
                                  extra={"torch": torch, "F": F, "copy": copy, "all_img_test": 1, "utils": None,
                                         "DTOffsetHelper": ns.DTOffsetHelper, "print": lambda *a, **k: None})

    def model(tok):
        v = index[tok.key]
        return (mask_logits[v][None], point[v][None], dir_logits[v][None])
    times = [0]
    base = get_probmaps(_Tok(), model, opt, "tile", times)
    g = {"np": np, "Image": Image, "img": _Tok(), "test_transform": lambda t: (t[0],), "get_probmaps": get_probmaps,
         "model": model, "opt": opt, "name": "tile", "times": times, "print": lambda *a, **k: None,
         "prob_maps": base[0], "point_maps": base[1], "prob_dcm": base[2]}
    exec(compile(_read_block(ns.root, "test_dam.py", 314, 450), "test_dam.py:314-450", "exec"), g)
    dcm = np.stack([np.asarray(g[n])[0] for n in ("prob_dcm", "prob_dcm_hf", "prob_dcm_vf", "prob_dcm_hvf", "prob_dcm_r90",
                                                  "prob_dcm_r90_hf", "prob_dcm_r90_vf", "prob_dcm_r90_hvf")])
    return g["prob_maps"], g["point_maps"], dcm


def _direction_one_hot(ns, target_direction0, target, direction_classes):
    """Runs train_util_dam.py:123-142 verbatim (torch CPU tensors in, float tensor [B,C,H,W] out)."""
    import copy
    import torch
    opt = types.SimpleNamespace(model={"direction": 1}, direction_classes=direction_classes)
    g = {"np": np, "torch": torch, "copy": copy, "opt": opt, "direction_label": True,
         "target_direction0": target_direction0, "target": target}
    code = _read_block(ns.root, "train_util_dam.py", 123, 142)
    exec(compile(code, "train_util_dam.py:123-142", "exec"), g)
    return g["target_direction0"]


def _plain_postprocess(ns, prob_maps, min_area=20, radius=2, postproc=0, model_name="modelName",
                       multi_class=True):
    """Runs test.py:270-295 verbatim."""
    import skimage.morphology as morph
    from skimage import measure
    from scipy import ndimage as ndi
    opt = types.SimpleNamespace(
        model={"multi_class": multi_class, "modelName": model_name},
        post={"min_area": min_area, "radius": radius, "postproc": postproc},
        transform={"test": {}})
    g = {"np": np, "morph": morph, "measure": measure, "ndi": ndi, "opt": opt,
         "postproc_other": importlib.import_module("postproc_other"), "prob_maps": prob_maps,
         "print": lambda *a, **k: None}
    code = _read_block(ns.root, "test.py", 270, 295)
    exec(compile(code, "test.py:270-295", "exec"), g)
    return {"pred_labeled": g["pred_labeled"], "pred_inside": g["pred_inside"], "pred2": g["pred2"]}
