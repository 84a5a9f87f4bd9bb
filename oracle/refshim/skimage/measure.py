import numpy as np
from scipy import ndimage as ndi


def label(input, neighbors=None, background=None, return_num=False, connectivity=None):
    """Equal-value connected components, background 0, raster-first numbering, int64."""
    x = np.asarray(input)
    if connectivity is None:
        connectivity = x.ndim
    st = ndi.generate_binary_structure(x.ndim, connectivity)
    bg = 0 if background is None else background
    vals = np.unique(x)
    vals = vals[vals != bg]
    if vals.size <= 1:
        lab, n = ndi.label(x != bg, st)
        lab = lab.astype(np.int64)
    else:
        # general equal-value case (not reached on the binary hot path): label each value, then
        # renumber by first raster occurrence
        lab = np.zeros(x.shape, dtype=np.int64)
        n = 0
        for v in vals:
            lv, nv = ndi.label(x == v, st)
            lab[lv > 0] = lv[lv > 0] + n
            n += nv
        flat = lab.ravel()
        first = np.full(n + 1, flat.size, dtype=np.int64)
        np.minimum.at(first, flat, np.arange(flat.size))
        order = np.argsort(first[1:], kind="stable")
        remap = np.zeros(n + 1, dtype=np.int64)
        remap[order + 1] = np.arange(1, n + 1)
        lab = remap[lab]
    return (lab, n) if return_num else lab
