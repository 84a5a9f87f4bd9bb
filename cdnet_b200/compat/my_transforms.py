"""my_transforms.py of the reference: `LabelEncoding` without direction targets (:661-837)."""
from ..training import LabelEncoding  # noqa: F401
