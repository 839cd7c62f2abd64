"""CPU: the C-ABI library builds, loads and exports every symbol include/b2attack.h
declares; argument validation fails loudly without touching a GPU."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT


def header_symbols():
    src = open(os.path.join(ROOT, "include", "b2attack.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_surface():
    syms = header_symbols()
    for s in ("b2_pgd_update", "b2_patch_apply", "b2_patch_update", "b2_cost_volume_fwd", "b2_cost_volume_bwd",
              "b2_grid_sample3d_fwd", "b2_grid_sample2d_fwd", "b2_grid_sample_bwd", "b2_conv3d",
              "b2_groupnorm_fwd", "b2_groupnorm_bwd", "b2_roi_align_fwd", "b2_roi_align_bwd", "b2_version",
              "b2_last_error"):
        assert s in syms


def test_library_exports_every_declared_symbol(built_lib):
    from eval_driving_safety_b200 import _lib
    for s in header_symbols():
        assert hasattr(built_lib, s), s
        assert s in _lib.SIGNATURES, "ctypes signature missing for " + s
    assert set(_lib.SIGNATURES) == set(header_symbols())
    assert built_lib.b2_version() >= 100


def test_bad_arguments_return_error_codes(built_lib):
    from eval_driving_safety_b200 import _lib
    null = ctypes.POINTER(ctypes.c_void_p)()
    rc = built_lib.b2_pgd_update(null, null, null, null, 0, 1, 3, 16, 0.1, 0.1, 0, None, None, None, None, None)
    assert rc == 10001 and "n_sets" in _lib.last_error()
    rc = built_lib.b2_conv3d(None, None, None, 1, 64, 64, 8, 8, 8, 1, 0, 1, None)
    assert rc == 10001 and "null" in _lib.last_error()
    rc = built_lib.b2_cost_volume_fwd(None, None, None, None, 1, 32, 4, 4, 4, 1, None)
    assert rc == 10001


def test_ops_refuse_cpu_tensors(built_lib):
    from eval_driving_safety_b200 import ops, attack, dsgn
    x = torch.zeros(1, 3, 4, 4)
    with pytest.raises(RuntimeError, match="CUDA"):
        attack.pgd_step(x, x, x, 0.1, 0.1)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.build_cost_volume(torch.zeros(1, 4, 4, 4), torch.zeros(1, 4, 4, 4), torch.zeros(1, 2))
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.conv3d(torch.zeros(1, 4, 4, 4, 4), torch.zeros(4, 4, 3, 3, 3))
    m = dsgn.StereoNet(dsgn.tiny_cfg()).freeze()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 32, 64), torch.zeros(1, 3, 32, 64), None, None, None)


def test_missing_library_fails_loudly(monkeypatch):
    from eval_driving_safety_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libb2attack.so")
    with pytest.raises(RuntimeError, match="no CPU"):
        _lib.load()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "eval_driving_safety_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f


def test_conv_fusion_caps_is_a_pure_host_query(built_lib):
    """b2_conv3d_fusion_caps tells which kernel serves a shape and what its epilogue can fuse; it
    launches nothing, so it must answer on a machine without a GPU."""
    def caps(n, cin, cout, d, h, w, stride, mode):
        rows, aok = ctypes.c_int(-1), ctypes.c_int(-1)
        assert built_lib.b2_conv3d_fusion_caps(n, cin, cout, d, h, w, stride, mode, ctypes.byref(rows), ctypes.byref(aok)) == 0
        return rows.value, aok.value

    assert caps(1, 64, 64, 48, 96, 312, 1, 0) == (148, 1)      # stride-1 kernel, one CTA per SM
    assert caps(1, 96, 64, 192, 20, 304, 1, 0) == (148, 1)
    assert caps(1, 128, 128, 24, 48, 156, 1, 0) == (148, 1)    # two N tiles share a CTA row
    assert caps(1, 128, 64, 24, 48, 156, 2, 1) == (148, 1)     # transposed (DECONV) kernel
    assert caps(1, 64, 128, 48, 96, 312, 2, 0) == (0, 0)       # stride-2 CONV: generic kernel, nothing fused
    assert caps(2, 64, 64, 8, 16, 16, 1, 0) == (0, 1)          # N > 1: no statistics (a CTA row would mix samples)
    assert caps(1, 64, 96, 8, 16, 16, 1, 0) == (0, 1)          # Nt = 48: 16-channel tail has no statistics
    assert caps(1, 64, 64, 1, 16, 8, 1, 0) == (1, 1)           # a single tile: one CTA, one row
    assert caps(1, 48, 64, 8, 16, 16, 1, 0) == (0, 0)          # Cin not a multiple of 32: not served by tcgen05
    # the fused entry point validates its arguments before touching the device
    from eval_driving_safety_b200 import _lib
    rc = built_lib.b2_conv3d_fused(None, None, None, None, None, 1, None, None, 1, 64, 64, 8, 8, 8, 1, 0, None)
    assert rc != 0 and "stat_mode" in _lib.last_error()
