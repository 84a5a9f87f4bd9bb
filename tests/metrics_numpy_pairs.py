"""TEST-ONLY numpy stand-in for cdnet_label_pairs (csrc/metrics.cu): the same (key, count) contract, so that the
host epilogue of cdnet_b200/metrics.py can be checked on a machine without a GPU."""
import numpy as np


def pair_arrays(true, pred, shuffle_seed=0):
    t = np.asarray(true).astype(np.uint64).ravel()
    q = np.asarray(pred).astype(np.uint64).ravel()
    keys, counts = np.unique((t << np.uint64(32)) | q, return_counts=True)
    perm = np.random.default_rng(shuffle_seed).permutation(keys.size)   # the kernel's order is arbitrary
    return keys[perm], counts[perm].astype(np.int32)
