"""The C-ABI library loads and exports every symbol include/az_engine.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "az_engine.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(az_[a-z_0-9]+)\s*\(", text)))


@pytest.fixture(scope="module")
def built_lib():
    from alphazero_implementation_b200 import _lib

    path = _lib.build_library()
    return path, _lib


def test_header_declares_the_expected_surface():
    syms = _declared_symbols()
    for must in ("az_create", "az_destroy", "az_env_step", "az_select_leaves", "az_gather_leaves", "az_expand_backup",
                 "az_run_simulations", "az_root_stats", "az_sample_moves", "az_drain_episodes", "az_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol(built_lib):
    path, _lib = built_lib
    lib = ctypes.CDLL(path)
    declared = _declared_symbols()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in az_engine.h but not exported"
    assert set(_lib.SIGNATURES) == set(declared), "ctypes signature table and header disagree"
    assert _lib.load().az_abi_version() == 1


def test_create_fails_loudly_without_gpu(built_lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    _, _lib = built_lib
    lib = _lib.load()
    cfg = _lib.AzConfig(6, 7, 4, 4, 10, 0, 0, 0, 1.0)
    h = ctypes.c_void_p()
    rc = lib.az_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc != 0 and not h.value
    assert b"no CPU path" in lib.az_last_error(None) or b"CUDA" in lib.az_last_error(None)
    from alphazero_implementation_b200 import Engine

    with pytest.raises(RuntimeError):
        Engine(4, 10)


def test_sass_is_sm100a_only(built_lib):
    import subprocess

    path, _ = built_lib
    out = subprocess.run(["cuobjdump", "--list-elf", path], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out), out
