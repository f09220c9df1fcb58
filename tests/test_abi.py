"""CPU: the C-ABI library loads and exports every symbol include/*.h declares; without a
CUDA device the compute entry points fail loudly (there is no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "tropical_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tnb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    from tropical import _native
    handle = ctypes.CDLL(_native.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 30
    for name in names:
        assert hasattr(handle, name), f"{name} declared in the header but not exported"
    # and the Python binding covers exactly the declared interface
    assert sorted(_native.SIGNATURES) == names


def test_library_identifies_itself():
    from tropical import _native
    L = _native.lib()
    assert L.tnb_version() >= 100
    assert L.tnb_device_count() >= 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_no_cpu_fallback():
    from tropical import _native
    from tropical.stanford.model import Net
    with pytest.raises(_native.NativeError):
        _native.NativeNet(4, 2, 19, 2, 2.5, 3, 16, np.zeros(8, np.float32), np.zeros(8, np.float32),
                          np.linspace(0, 1, 5, dtype=np.float32))
    net = Net()
    with pytest.raises(_native.NativeError), torch.no_grad():   # the extraction path is no_grad
        net.sdf(torch.zeros(4, 3))
    import tropical.subpoly as sp
    with pytest.raises(_native.NativeError):
        sp.subpoly(net, 3, 1.2, force=True)
    with pytest.raises(_native.NativeError):
        sp.subpoly_batch([net, net], 3, 1.2, force=True)
    from tropical import subpoly_debug as dbg
    z = torch.zeros
    with pytest.raises(_native.NativeError):   # an edge off its planes: the repair has no CPU route either
        dbg.deal_with_gradient_descent(torch.ones(1, dtype=torch.bool), torch.ones(1, 2), z(1, 2, 3), 1e-4, z(1, dtype=torch.bool), 5,
                                       z(1, 2, dtype=torch.long), torch.full((1, 3), 0.5), net)


def test_net_create_validates_arguments():
    from tropical import _native
    L = _native.lib()
    d = _native.NetDesc(4, 3, 19, 2, 2.5, 3, 16, 1.0, 1e-4, None, 0, None, 0, None, 0)
    h = ctypes.c_void_p()
    assert L.tnb_net_create(ctypes.byref(d), ctypes.byref(h)) == -4  # n_features != 2
    assert b"n_features" in L.tnb_last_error()
    d.n_features, d.num_hidden = 2, 1000
    assert L.tnb_net_create(ctypes.byref(d), ctypes.byref(h)) == -1
