"""Stand-in for the subset of skimage.feature the reference calls (TEST INFRASTRUCTURE ONLY): peak_local_max as
my_transforms.py:775 uses it -- `peak_local_max(distance_i, exclude_border=0, num_peaks=1)`.

Restated from the public description of scikit-image 0.16-0.18 (no copy of scikit-image exists here): candidates are the
pixels that equal the maximum of their (2*min_distance+1)^2 neighbourhood and exceed the threshold
(max(threshold_abs or image.min(), threshold_rel * image.max())); they are returned highest intensity first.
PARITY UNPINNED: among candidates of EQUAL intensity scikit-image's order comes from numpy's default (unstable) argsort
of the intensities; this restatement breaks such ties in raster order (what a stable sort -- numpy's insertion sort below
16 candidates -- gives)."""
import numpy as np
from scipy import ndimage as ndi


def peak_local_max(image, min_distance=1, threshold_abs=None, threshold_rel=None, exclude_border=True, indices=True,
                   num_peaks=np.inf, footprint=None, labels=None, num_peaks_per_label=np.inf):
    image = np.asarray(image)
    if labels is not None or footprint is not None:
        raise NotImplementedError("refshim: peak_local_max with labels / footprint")
    if image.size == 0 or np.all(image == image.flat[0]):
        return np.empty((0, image.ndim), dtype=np.intp) if indices else np.zeros(image.shape, bool)
    size = 2 * int(min_distance) + 1
    mx = ndi.maximum_filter(image, size=size, mode="constant")
    mask = image == mx
    border = int(min_distance) if exclude_border is True else int(exclude_border or 0)
    if border:
        for ax in range(image.ndim):
            sl = [slice(None)] * image.ndim
            sl[ax] = slice(0, border)
            mask[tuple(sl)] = False
            sl[ax] = slice(-border, None)
            mask[tuple(sl)] = False
    thr_abs = image.min() if threshold_abs is None else threshold_abs
    thr = max(thr_abs, (0.0 if threshold_rel is None else threshold_rel) * image.max())
    mask &= image > thr
    coord = np.argwhere(mask)  # raster order
    order = np.argsort(-image[mask], kind="stable")
    coord = coord[order]
    if np.isfinite(num_peaks) and len(coord) > num_peaks:
        coord = coord[:int(num_peaks)]
    if indices:
        return coord
    out = np.zeros(image.shape, bool)
    out[tuple(coord.T)] = True
    return out
