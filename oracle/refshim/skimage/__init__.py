"""Stand-in for the scikit-image subset used by CDNet's geometry path (TEST INFRASTRUCTURE)."""
from . import morphology, measure, segmentation, io, color, feature, filters  # noqa: F401

__version__ = "0.0-cdnet-oracle-shim"
