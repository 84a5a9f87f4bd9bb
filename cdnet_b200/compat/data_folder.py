"""Label loading of the reference's dataset class without the uint8 truncation (data_folder.py:20-41).

`img_loader(path, num_channels)` keeps the reference's behaviour bit for bit (instance maps from `.mat` / `.npy` files go
through `astype(np.uint8)`, data_folder.py:26,29,37: ids wrap at 256, so a tile with more than 255 nuclei re-uses ids and
two touching nuclei with ids congruent mod 256 lose the boundary between them).  With `keep_ids=True` the instance map is
returned as an int32 array of the original ids instead; cdnet_b200.api.LabelEncoding takes such an image through the
int32 entry point of the target transform (cdnet_encode_targets_i32).  Host-side file I/O only -- nothing here runs on
the GPU."""
import numpy as np


def img_loader(path, num_channels, keep_ids=False):
    from PIL import Image
    is_mat, is_npy = ".mat" in path, ".npy" in path
    if is_mat or is_npy:
        if is_mat:
            import scipy.io as scio
            img = scio.loadmat(path)["inst_map"]
        else:
            img = np.load(path)
        if keep_ids:
            a = np.asarray(img)
            if a.size and (a.min() < 0 or a.max() > 2 ** 31 - 1):
                raise ValueError("instance ids must fit int32")
            return np.ascontiguousarray(a, dtype=np.int32)
        if is_mat and num_channels != 1:
            return img  # data_folder.py:32-33 hands the raw array on for multi-channel .mat labels
        return Image.fromarray(np.asarray(img).astype(np.uint8))  # data_folder.py:26,29,37
    if num_channels == 1:
        return Image.open(path)
    return Image.open(path).convert("RGB")
