"""Import-only stand-in so the reference's modules load here (no plotting is used)."""
