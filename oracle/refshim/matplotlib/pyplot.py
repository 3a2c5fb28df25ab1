"""Empty pyplot stand-in."""
