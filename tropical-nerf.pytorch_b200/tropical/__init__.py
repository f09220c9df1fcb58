"""`tropical` -- host-side mirror of the reference package for the mesh-extraction path.

Same public names as the reference's `tropical/__init__.py` (TropicalHashGrid, Tropical,
torch.ext helpers, `deprecated`); the work runs in hand-written sm_100a kernels behind
the C ABI of `include/tropical_b200.h` (see `_native.py`).  No CPU fallback.
"""
import functools
import warnings

from .tropical import *  # noqa: F401,F403
from . import torch_ext

import torch
torch.ext = torch_ext  # the reference exposes its helpers as torch.ext (tropical/__init__.py:9)


def deprecated(reason=None):
    """Mark a function as deprecated (tropical/__init__.py:12-34)."""
    def decorator(func):
        @functools.wraps(func)
        def wrapped(*args, **kwargs):
            message = f"Function '{func.__name__}' is deprecated."
            if reason:
                message += f" Reason: {reason}"
            warnings.warn(message, category=DeprecationWarning, stacklevel=2)
            return func(*args, **kwargs)
        return wrapped
    return decorator
