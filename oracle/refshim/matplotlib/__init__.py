"""Empty stub so the reference modules import in the dev container (oracle/make_golden.py only)."""
