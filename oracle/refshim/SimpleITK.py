"""import-only stub (image I/O library, not on the hot path)."""
