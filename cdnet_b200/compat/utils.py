"""utils.py of the reference: `DcmVoting2` (:1150-1159)."""
from ..api import DcmVoting2  # noqa: F401
