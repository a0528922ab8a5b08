"""No-GPU checks of the drop-in boundary: the C-ABI library builds, loads and exports exactly the entry
points include/stratego_b200.h declares; the host layer refuses to run without a CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "stratego_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sx_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from stratego_env_b200 import _lib
    names = _declared()
    assert len(names) >= 20
    lib = _lib.load()
    raw = ctypes.CDLL(_lib.library_path())
    for name in names:
        assert hasattr(raw, name), "header declares %s but the library does not export it" % name
    assert sorted(_lib.SYMBOLS) == names, "ctypes binding and header are out of sync"
    assert lib.sx_version() >= 1


def test_config_validation_without_gpu():
    """sx_config_create is pure host code: geometry, strides and error behaviour (penv:28-30)"""
    from stratego_env_b200 import _lib
    lib = _lib.load()
    desc = _lib.SxConfigDesc()
    obst = np.zeros(100, np.uint8)
    cap = np.zeros((12, 9), np.float32)
    rec, unit = np.zeros(5, np.float32), np.zeros(2, np.float32)
    desc.rows, desc.cols, desc.max_turns, desc.usable_rows = 10, 10, 1000, 4
    for code, n in {1: 1, 2: 2, 3: 1, 9: 1, 10: 1, 11: 1, 12: 1}.items():
        desc.piece_amounts[code] = n
    desc.obstacles, desc.captured_lut = obst.ctypes.data, cap.ctypes.data
    desc.recent_lut, desc.unit_lut = rec.ctypes.data, unit.ctypes.data
    handle = ctypes.c_void_p()
    assert lib.sx_config_create(ctypes.byref(desc), ctypes.byref(handle)) == 0
    lay = _lib.SxLayout()
    assert lib.sx_config_layout(handle, ctypes.byref(lay)) == 0
    assert (lay.rows, lay.cols, lay.cells, lay.spatial_channels, lay.spatial_actions, lay.action_size) == \
        (10, 10, 100, 37, 3700, 2001)                     # impl:253-259
    assert (lay.po_floats, lay.fo_floats, lay.pieces_per_side, lay.setup_len) == (6700, 7900, 8, 40)
    assert lay.board_stride % 16 == 0 and lay.captured_stride >= 16
    lib.sx_config_destroy(handle)
    desc.rows = 2
    assert lib.sx_config_create(ctypes.byref(desc), ctypes.byref(handle)) != 0
    assert b"at least 3" in lib.sx_last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from stratego_env_b200 import BatchedStrategoEnv, GameVersions, StrategoMultiAgentEnv, StrategoProceduralEnv
    from stratego_env_b200._lib import StrategoB200Error
    with pytest.raises(StrategoB200Error):
        StrategoMultiAgentEnv({"version": GameVersions.TINY})
    with pytest.raises(StrategoB200Error):
        BatchedStrategoEnv({"version": GameVersions.TINY}, num_envs=4)
    with pytest.raises(StrategoB200Error):
        StrategoProceduralEnv(4, 4)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "stratego_env_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "stratego_oracle" not in text, f
