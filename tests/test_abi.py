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


def _sass_of(path, pattern):
    """Instructions [(address, text)] of the first kernel whose mangled name matches `pattern`."""
    import subprocess

    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    ins, inside = [], False
    for line in out.splitlines():
        if "Function :" in line:
            if inside:
                break
            inside = re.search(pattern, line) is not None
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?)\s*;", line) if inside else None
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    return ins


def test_level_loop_of_the_fused_tree_kernel_stays_straight_line(built_lib):
    """The descent's level loop is a dependent chain the kernel's speed hangs on: it must stay free of
    reconvergence regions (an `if` that the compiler turns into a divergent branch), read its tables with LDS (tables in
    shared memory are a compile-time fact) and use the tensor-core-free fp64 path only.  Checked on the SASS of the variant
    the bench runs (4 trees per warp, uniform evaluator, latency variant, fused move, shared-memory tables)."""
    path, _ = built_lib
    ins = _sass_of(path, r"k_run_simsILi4ELi1ELb1ELb1ELb1E")
    assert len(ins) > 1000
    loops = []
    for i, (addr, text) in enumerate(ins):
        m = re.search(r"\bBRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", text)
        if m and int(m.group(1), 16) < addr:
            body = [t for a, t in ins if int(m.group(1), 16) <= a <= addr]
            if any("SHFL.BFLY" in t for t in body):
                loops.append(body)
    level = min(loops, key=len)  # innermost loop with the arg-max butterfly = one tree level
    assert 100 < len(level) <= 170, len(level)
    assert not any(t.startswith(("BSSY", "BSYNC")) or " BSSY" in t for t in level)
    assert sum("SHFL.BFLY" in t for t in level) == 9 and sum("SHFL.IDX" in t for t in level) == 3
    assert any("LDS.128" in t for t in level) and not any(re.search(r"\bLD\.E", t) for t in level)
    # two divisions (multiply + 4 FMAs each), their sum, and the next level's numerator (c * P) * sqrt(N): 13 fp64 operations
    assert sum(t.split()[0 if not t.startswith("@") else 1].startswith(("DFMA", "DMUL", "DADD")) for t in level) <= 13


def test_rules_kernels_split_their_integer_work_over_both_issue_pipes(built_lib):
    """The streaming rules kernels are bound by integer issue, not by HBM: what made them faster is the split
    of that work over the ALU pipe (logic ops, funnel shifts, selects) and the FMA pipe (shifts written as multiplications).  Checked
    on the SASS of the variants the library runs by default: straight-line code, and per 4 positions at most 440 / 215 ALU-pipe
    instructions (the 64-bit kernels of round 1: 697 / 379) beside at least 250 / 170 integer multiply-adds."""
    path, _ = built_lib
    alu = ("LOP3", "SHF", "IADD3", "ISETP", "SEL", "PRMT", "PLOP3", "LEA", "VIADD", "P2R", "R2P")
    for pattern, max_alu, min_fma in ((r"k_env_step_hILi4E", 440, 250), (r"k_state_info_hILi4E", 215, 170)):
        ins = [t.split()[1] if t.startswith("@") else t.split()[0] for _, t in _sass_of(path, pattern)]
        assert len(ins) > 300, pattern
        assert not any(op.startswith(("BSSY", "BSYNC", "CALL")) for op in ins)
        n_alu = sum(op.startswith(alu) for op in ins)
        n_fma = sum(op.startswith("IMAD") for op in ins)
        assert n_alu <= max_alu and n_fma >= min_fma, (pattern, n_alu, n_fma)
