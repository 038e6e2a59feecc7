"""NumPy-backed stand-in for the `mlx` package (oracle/test infrastructure only)."""
