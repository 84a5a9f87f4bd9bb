"""cdnet_b200 -- B200-native (sm_100a) implementation of CDNet's geometry hot path.

Drop-in callables with the reference's signatures live in `cdnet_b200.api` (re-exported here; inference
post-processing and the direction-aware target transform), `cdnet_b200.training` (DTOffsetHelper, the
direction one-hot block, my_transforms.LabelEncoding), `cdnet_b200.metrics` (stats_utils.py) and
`cdnet_b200.sharded` (whole-slide row partition); they call hand-written CUDA kernels through the C ABI of
libcdnet_b200.so.  There is no CPU
fallback: without the library or without a Blackwell GPU the calls raise.
"""
from ._cabi import CdnetError  # noqa: F401


def __getattr__(name):
    # api imports torch; keep `import cdnet_b200` cheap and let the attribute access pull it in
    if name.startswith("__"):
        raise AttributeError(name)
    import importlib
    api = importlib.import_module(".api", __name__)
    try:
        return getattr(api, name)
    except AttributeError:
        raise AttributeError("module 'cdnet_b200' has no attribute %r" % name)
