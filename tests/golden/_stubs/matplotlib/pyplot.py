"""Import-only stand-in (`tropical/geometry.py:10` imports pyplot but never calls it on the path)."""
