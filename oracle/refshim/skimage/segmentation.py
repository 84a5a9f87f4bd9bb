import numpy as np


def watershed(image, markers=None, connectivity=1, offset=None, mask=None, compactness=0,
              watershed_line=False):
    """Marker-controlled watershed restated in oracle/ws_flood.c (stable (value, age, index)
    order -- the canonical order of this build; "parity unpinned" against real scikit-image)."""
    from oracle import clib
    assert connectivity == 1 and offset is None and compactness == 0 and not watershed_line
    image = np.asarray(image)
    assert image.ndim == 2 and markers is not None
    if mask is None:
        mask = np.ones(image.shape, dtype=bool)
    mask = np.asarray(mask)
    markers = np.asanyarray(markers) * mask
    return clib.ws_flood(image.astype(np.float64), markers.astype(np.int32),
                         mask.astype(np.int8) != 0, order=_ORDER[0])


_ORDER = ["stable"]
