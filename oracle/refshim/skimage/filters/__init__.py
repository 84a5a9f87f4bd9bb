"""import-only stub (not on the hot path)."""
from . import rank  # noqa: F401
