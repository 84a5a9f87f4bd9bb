"""postproc_other.py of the reference: `process` (postproc_other.py:15-54)."""
from ..api import process  # noqa: F401
