"""CPU-side checks of the native library: it builds for sm_100a, loads, and exports every symbol that
include/tdb200.h declares (no compute calls without a GPU)."""
import os
import re

from torch_de_solver_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = _native.load()
    header = open(os.path.join(ROOT, 'include', 'tdb200.h')).read()
    declared = set(re.findall(r'\b(tdb200_[a-z_0-9]+)\s*\(', header))
    assert declared, 'no prototypes found'
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, missing
    assert declared == set(_native.EXPORTS)
    assert lib.tdb200_version() >= 100


def test_plan_create_without_gpu_fails_loudly():
    import ctypes as C
    import torch
    if torch.cuda.is_available():
        return
    lib = _native.load()
    net = _native.NetDesc()
    net.n_layers = 2
    net.widths[0], net.widths[1], net.widths[2] = 2, 8, 1
    import numpy as np
    from torch_de_solver_b200.plan import SEGMENT_DTYPE
    seg = np.zeros(1, SEGMENT_DTYPE)
    seg['K'] = seg['M'] = seg['n_cols'] = seg['identity'] = 1
    h = C.c_void_p()
    rc = lib.tdb200_plan_create(C.byref(net), 1, _native.np_ptr(seg), 0, None, 0, None, 0, None, 1, 0, C.byref(h))
    assert rc < 0
    assert b'CUDA' in lib.tdb200_last_error() or b'device' in lib.tdb200_last_error()


def test_sass_is_sm100():
    import shutil
    import subprocess
    if shutil.which('cuobjdump') is None:
        return
    out = subprocess.run(['cuobjdump', '-lelf', _native.lib_path()], capture_output=True, text=True).stdout
    assert 'sm_100a' in out
