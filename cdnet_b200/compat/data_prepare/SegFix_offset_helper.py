"""data_prepare/SegFix_offset_helper.py of the reference: DTOffsetConfig (:21-46), Sobel (:97-132), DTOffsetHelper
(:246-261, 286-341, 423-506)."""
from ...training import DTOffsetConfig, DTOffsetHelper, Sobel  # noqa: F401
