"""import-only stub."""
