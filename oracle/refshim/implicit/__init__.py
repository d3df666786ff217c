"""Empty import-time stub (oracle/make_golden.py only)."""
