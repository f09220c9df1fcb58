"""Import-only stand-in (`tropical/utils/chamfer_distance.py:12`); never called on the path."""
