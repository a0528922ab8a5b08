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


def test_binding_signatures_match_the_header():
    """every prototype of include/stratego_b200.h has as many parameters as the ctypes binding passes, and the struct
    bindings have the header's field counts (a drifted signature corrupts the stack silently)"""
    import re
    from stratego_env_b200 import _lib
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "stratego_b200.h")).read(), flags=re.S)
    protos = re.findall(r"^(?:const\s+)?[A-Za-z_][\w\s\*]*?\b(sx_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.M)
    seen = {}
    for name, params in protos:
        params = params.strip()
        seen[name] = 0 if params in ("", "void") else params.count(",") + 1
    for name, (restype, argtypes) in _lib.SYMBOLS.items():
        assert name in seen, name
        assert seen[name] == len(argtypes), (name, seen[name], len(argtypes))

    def fields(struct_name):
        body = re.search(r"typedef struct\s*\{([^{}]*)\}\s*%s\s*;" % struct_name, text).group(1)
        n = 0
        for decl in body.split(";"):
            decl = re.sub(r"\[[^\]]*\]", "", decl).strip()
            if decl:
                n += decl.count(",") + 1
        return n
    assert fields("sx_outputs") == len(_lib.SxOutputs._fields_)
    assert fields("sx_state") == len(_lib.SxState._fields_)
    assert fields("sx_layout") == len(_lib.SxLayout._fields_)
    assert fields("sx_config_desc") == len(_lib.SxConfigDesc._fields_)


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
    # launch tuning is host state only: accepted ranges, -1 keeps, null config rejected
    assert lib.sx_config_set_tuning(handle, 10, 2, 1) == 0
    assert lib.sx_config_set_tuning(handle, -1, -1, -1) == 0
    assert lib.sx_config_set_tuning(handle, 33, 0, 0) != 0 and b"warps_per_block" in lib.sx_last_error()
    assert lib.sx_config_set_tuning(handle, 8, 3, 0) != 0
    assert lib.sx_config_set_tuning(None, 8, 1, 0) != 0
    # curriculum table: host state only; a table needs all three tensors, n_states = 0 switches back to setups
    assert lib.sx_config_set_start_states(handle, _lib.SxState(), 0, None, 0, 0) == 0
    partial = _lib.SxState()
    partial.board = 4096
    assert lib.sx_config_set_start_states(handle, partial, 3, None, 0, 0) != 0 and b"board, aux and captured" in lib.sx_last_error()
    assert lib.sx_config_set_start_states(handle, partial, -1, None, 0, 0) != 0
    assert lib.sx_config_set_start_states(None, _lib.SxState(), 0, None, 0, 0) != 0
    lib.sx_config_destroy(handle)
    desc.rows = 2
    assert lib.sx_config_create(ctypes.byref(desc), ctypes.byref(handle)) != 0
    assert b"at least 3" in lib.sx_last_error()


def test_original_channel_mode_layout_without_gpu():
    """obs_channel_mode = SX_CHANNELS_ORIGINAL: 32 / 33 channels (impl:1148, impl:1070), and the two extra tables are
    required"""
    from stratego_env_b200 import _lib
    from stratego_env_b200.config import (MICRO_STRATEGO_CONFIG as CFG, original_captured_lut, original_po_rank_lut,
                                          original_rank_lut, original_unit_lut, piece_amounts_array, recent_moves_lut)
    lib = _lib.load()
    desc = _lib.SxConfigDesc()
    desc.rows, desc.cols, desc.max_turns, desc.usable_rows = CFG["rows"], CFG["columns"], CFG["max_turns"], 1
    for code, n in enumerate(piece_amounts_array(CFG["piece_amounts"])):
        desc.piece_amounts[code] = int(n)
    keep = [np.zeros(12, np.uint8), original_captured_lut(CFG["piece_amounts"]), recent_moves_lut(), original_unit_lut(),
            original_rank_lut(), original_po_rank_lut()]
    desc.obstacles, desc.captured_lut, desc.recent_lut, desc.unit_lut = (k.ctypes.data for k in keep[:4])
    desc.obs_channel_mode = _lib.SX_CHANNELS_ORIGINAL
    handle = ctypes.c_void_p()
    assert lib.sx_config_create(ctypes.byref(desc), ctypes.byref(handle)) != 0   # rank tables missing
    assert b"rank_lut" in lib.sx_last_error()
    desc.rank_lut, desc.po_rank_lut = keep[4].ctypes.data, keep[5].ctypes.data
    assert lib.sx_config_create(ctypes.byref(desc), ctypes.byref(handle)) == 0
    lay = _lib.SxLayout()
    assert lib.sx_config_layout(handle, ctypes.byref(lay)) == 0
    assert (lay.po_channels, lay.fo_channels, lay.po_floats, lay.fo_floats) == (32, 33, 12 * 32, 12 * 33)
    lib.sx_config_destroy(handle)
    desc.obs_channel_mode = 7
    assert lib.sx_config_create(ctypes.byref(desc), ctypes.byref(handle)) != 0
    # the normalised tables themselves: rank / 12 and PO rank / 13 mapped to [-1, 1] (maenv:87-199, 388-396, 499-508)
    assert keep[4][0] == -1.0 and keep[4][12] == 1.0 and keep[5][13] == 1.0 and keep[3].tolist() == [-1.0, 0.0]


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


def test_policy_uniform_is_strictly_inside_the_unit_interval():
    """sx_policy.cu uniform_open01: (k + 0.5) / 2^23 for the 23 top bits k of a Philox word, in float32 arithmetic.
    Both extremes stay strictly inside (0, 1), so the Gumbel score -log(-log(u)) is finite (the previous 24-bit form
    rounded to exactly 1.0 for the top value)."""
    import numpy as np
    for word in (0x00000000, 0x000001FF, 0xFFFFFFFF, 0xFFFFFE00, 0x80000000):
        k = np.float32(word >> 9)
        u = (k + np.float32(0.5)) * np.float32(1.0 / 8388608.0)
        assert u.dtype == np.float32 and np.float32(0.0) < u < np.float32(1.0), (hex(word), u)
        assert np.isfinite(-np.log(-np.log(u)))
    old = (np.float32(0xFFFFFFFF >> 8) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)
    assert old == np.float32(1.0)  # what the fix removes
