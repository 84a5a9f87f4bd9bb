"""skimage.morphology subset restated on scipy.ndimage (see ../../README.md)."""
import numpy as np
from scipy import ndimage as ndi

from . import selem  # noqa: F401
from .selem import disk  # noqa: F401


def _default_selem(image):
    # documented default: "cross-shaped structuring element (connectivity=1)"
    return ndi.generate_binary_structure(image.ndim, 1)


def dilation(image, selem=None, out=None):
    image = np.asarray(image)
    fp = _default_selem(image) if selem is None else np.asarray(selem)
    fp = fp[tuple(slice(None, None, -1) for _ in range(fp.ndim))]  # skimage un-inverts scipy's flip
    if out is None:
        out = np.empty_like(image)
    ndi.grey_dilation(image, footprint=fp, output=out)
    return out


def erosion(image, selem=None, out=None):
    image = np.asarray(image)
    fp = _default_selem(image) if selem is None else np.asarray(selem)
    if out is None:
        out = np.empty_like(image)
    ndi.grey_erosion(image, footprint=fp, output=out)
    return out


def remove_small_objects(ar, min_size=64, connectivity=1, in_place=False):
    ar = np.asarray(ar)
    if not (ar.dtype == bool or np.issubdtype(ar.dtype, np.integer)):
        raise TypeError("Only bool or integer image types are supported. Got %s." % ar.dtype)
    out = ar if in_place else ar.copy()
    if min_size == 0:
        return out
    if out.dtype == bool:
        st = ndi.generate_binary_structure(ar.ndim, connectivity)
        ccs = np.zeros_like(ar, dtype=np.int32)
        ndi.label(ar, st, output=ccs)
    else:
        ccs = out
    try:
        sizes = np.bincount(ccs.ravel())
    except ValueError:
        raise ValueError("Negative value labels are not supported.")
    too_small = sizes < min_size
    out[too_small[ccs]] = 0
    return out


def watershed(image, markers=None, connectivity=1, offset=None, mask=None, compactness=0,
              watershed_line=False):
    from ..segmentation import watershed as _ws
    return _ws(image, markers, connectivity, offset, mask, compactness, watershed_line)
