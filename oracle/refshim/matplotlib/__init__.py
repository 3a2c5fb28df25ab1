"""Empty matplotlib stand-in (reference density.py imports pyplot at module scope)."""
