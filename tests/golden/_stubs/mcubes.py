"""Import-only stand-in; never called on the path."""
