"""import-only stub (augmentation library, not on the hot path)."""
