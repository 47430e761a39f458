"""The C-ABI library loads on a CPU-only box and exports every symbol include/ilqg_b200.h declares."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def declared_functions():
    txt = open(os.path.join(ROOT, "include", "ilqg_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ilqg_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_the_hot_path():
    names = declared_functions()
    for must in ("ilqg_create", "ilqg_destroy", "ilqg_fd_batch_dev", "ilqg_fd_batch_host", "ilqg_step_batch_host",
                 "ilqg_forward_batch_host", "ilqg_compile_mjcf"):
        assert must in names


def test_library_exports_every_declared_symbol(pkg):
    L = pkg.lib()
    missing = [n for n in declared_functions() if not hasattr(L, n)]
    assert not missing, missing


def test_model_struct_size_matches_fixture(pkg):
    n = pkg.lib().ilqg_model_sizeof()
    for name in ("inverted_pendulum", "hopper", "humanoid"):
        assert os.path.getsize(os.path.join(pkg.MODELS_DIR, name + ".ilqgm")) == n


def test_create_fails_loudly_without_gpu_or_on_bad_model(pkg):
    import torch
    bad = pkg.Model.named("hopper").copy()
    bad.buf[0] ^= 0xFF  # corrupt the magic
    h = C.c_void_p()
    rc = pkg.lib().ilqg_create(bad.ptr, 0, C.byref(h))
    assert rc == pkg.ERR_MODEL
    if not torch.cuda.is_available():
        with pytest.raises(pkg.IlqgError) as e:
            pkg.Handle(pkg.Model.named("hopper"), 0)
        assert e.value.code == pkg.ERR_CUDA  # no CPU fallback behind the ABI


def test_null_arguments_are_rejected(pkg):
    L = pkg.lib()
    assert L.ilqg_create(None, 0, None) == pkg.ERR_ARG
    assert L.ilqg_fd_batch_host(None, 1, None, None, None, None, None, None, None, None, None) == pkg.ERR_ARG
    assert L.ilqg_destroy(None) == pkg.OK
    assert L.ilqg_deriv_size(pkg.Model.named("hopper").ptr) == 105
    assert L.ilqg_deriv_size(pkg.Model.named("inverted_pendulum").ptr) == 15
    assert L.ilqg_deriv_size(pkg.Model.named("humanoid").ptr) == 2100


def test_model_tables_are_validated(pkg, tmp_path):
    """A table that did not come out of the compiler (a file, a buffer over the ABI) is checked before the kernels index with it."""
    L = pkg.lib()
    err = C.create_string_buffer(256)
    for name in ("inverted_pendulum", "hopper", "humanoid"):
        assert L.ilqg_model_validate(pkg.Model.named(name).ptr, err, 256) == pkg.OK, err.value
    # out-of-range indices and counts
    for field, idx, val, code in (("body_parentid", 2, 9, pkg.ERR_MODEL), ("dof_parentid", 1, 5, pkg.ERR_MODEL), ("pair_geom2", 0, 23, pkg.ERR_MODEL),
                                  ("act_dofid", 0, 31, pkg.ERR_MODEL), ("jnt_type", 3, 1, pkg.ERR_UNSUPPORTED),   # a LIMITED ball joint (and the dof counts no longer add up)
                                  ("geom_type", 1, 5, pkg.ERR_UNSUPPORTED), ("npair", 0, 4000, pkg.ERR_MODEL)):
        bad = pkg.Model.named("hopper").copy()
        bad.field(field)[idx] = val
        rc = L.ilqg_model_validate(bad.ptr, err, 256)
        assert rc == code, (field, rc, err.value)
        assert field.split("_")[0].encode() in err.value or b"count" in err.value or b"joint" in err.value or b"geom" in err.value
        h = C.c_void_p()
        assert L.ilqg_create(bad.ptr, 0, C.byref(h)) == code    # refused before any CUDA call
        with pytest.raises(pkg.IlqgError):
            bad.validate()
    # a file of the wrong size is not a table
    short = tmp_path / "short.ilqgm"
    short.write_bytes(pkg.Model.named("hopper").buf.tobytes()[:-8])
    buf = np.zeros(L.ilqg_model_sizeof(), np.uint8)
    assert L.ilqg_model_load(str(short).encode(), buf.ctypes.data_as(C.c_void_p)) == pkg.ERR_IO
    with pytest.raises(ValueError):
        pkg.Model(np.zeros(100, np.uint8))
