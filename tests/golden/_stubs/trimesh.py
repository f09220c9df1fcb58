"""Import-only stand-in (`tropical/utils/chamfer_distance.py:7`); never called on the path."""
