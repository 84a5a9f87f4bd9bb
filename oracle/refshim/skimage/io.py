"""import-only stub (not on the hot path)."""
