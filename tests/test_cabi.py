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
