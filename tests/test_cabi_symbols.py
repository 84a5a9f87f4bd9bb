"""CPU: the C-ABI library loads and exports every symbol include/cdnet_b200.h declares (no compute)."""
import ctypes
import os
import re

from conftest import REPO


def _declared():
    src = open(os.path.join(REPO, "include", "cdnet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cdnet_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from cdnet_b200 import _cabi
    L = _cabi.lib()
    names = _declared()
    assert len(names) >= 25
    raw = ctypes.CDLL(_cabi.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), "libcdnet_b200.so does not export %s" % n
    # the ctypes table and the header agree
    assert set(_cabi.SIGNATURES) == set(names), set(_cabi.SIGNATURES) ^ set(names)
    assert L.cdnet_version().startswith(b"cdnet_b200")


def test_workspace_queries_need_no_gpu():
    from cdnet_b200 import _cabi
    L = _cabi.lib()
    n = 14 * 1000 * 1000
    assert L.cdnet_ddm_workspace_bytes(14, 1000, 1000) >= 2 * n
    assert L.cdnet_dam_postproc_workspace_bytes(14, 1000, 1000) >= 20 * n
    assert L.cdnet_encode_targets_workspace_bytes(2, 500, 500) > 0
    assert L.cdnet_ccl_workspace_bytes(0, 10, 10) == 0  # invalid arguments -> 0
    assert L.cdnet_ws_postproc_workspace_bytes(1, 70000, 70000) == 0  # H*W must stay below 2^31


def test_product_does_not_import_oracle():
    pkg = os.path.join(REPO, "cdnet_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(root, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
                # the SIMT emulator (tests/simt) is test infrastructure too: the product must not load it
                assert "import simt" not in txt and "from simt" not in txt and "_simt.so" not in txt, f


def test_no_cpu_fallback_without_cuda():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import numpy as np
    from cdnet_b200 import api, CdnetError
    with pytest.raises(CdnetError):
        api.generate_dd_map(np.zeros((4, 4), np.uint8), 9)


def test_compat_modules_mirror_the_reference_names():
    """cdnet_b200.compat.* carry the reference's module and callable names (import-line drop-in)"""
    import importlib
    want = {"postproc_other": ["process"],
            "data_prepare.getDirectionDiffMap": ["generate_dd_map", "circshift"],
            "data_prepare.SegFix_offset_helper": ["DTOffsetHelper", "DTOffsetConfig", "Sobel"],
            "my_transforms_direction": ["LabelEncoding", "get_centerpoint2"],
            "my_transforms": ["LabelEncoding"],
            "stats_utils": ["get_fast_aji", "get_fast_aji_plus", "get_fast_pq", "get_dice_1", "get_dice_2",
                            "get_fast_dice_2", "remap_label"],
            "utils": ["DcmVoting2"]}
    for mod, names in want.items():
        m = importlib.import_module("cdnet_b200.compat." + mod)
        for n in names:
            assert callable(getattr(m, n)), (mod, n)
    from cdnet_b200 import api, training
    from cdnet_b200.compat import my_transforms, my_transforms_direction
    assert my_transforms_direction.LabelEncoding is api.LabelEncoding
    assert my_transforms.LabelEncoding is training.LabelEncoding
    for n in ("label_to_vector", "align_angle", "angle_to_vector", "vector_to_label", "angle_to_direction_label"):
        assert callable(getattr(training.DTOffsetHelper, n))
