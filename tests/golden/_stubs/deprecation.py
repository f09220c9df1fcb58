"""Import-only stand-in for the `deprecation` package (`tropical/geometry.py:7`)."""


def deprecated(*args, **kwargs):
    def deco(fn):
        return fn
    return deco
