"""TEST INFRASTRUCTURE — ctypes wrapper + build recipe for `c4_oracle.c`.

Only tests/, `__graft_entry__` (build / smoke) and bench.py's CPU-baseline legs
import this.  See the header of c4_oracle.c for what is restated and from where.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "c4_oracle.c")
LIB = os.path.join(HERE, "libc4oracle.so")

EVAL_CB = C.CFUNCTYPE(
    None, C.c_void_p, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint8),
    C.POINTER(C.c_uint8), C.POINTER(C.c_float), C.POINTER(C.c_float),
)


def build(force: bool = False) -> str:
    """gcc -O2, no FMA contraction (fp64 ops must round one by one like CPython's)."""
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", "-std=gnu11", "-o", LIB, SRC, "-lm"]
    subprocess.run(cmd, check=True, cwd=HERE)
    return LIB


class _Cfg(C.Structure):
    _fields_ = [
        ("E", C.c_int32), ("S", C.c_int32), ("eval_kind", C.c_int32), ("init_player", C.c_int32),
        ("quota", C.c_int32), ("max_steps", C.c_int32), ("c_puct", C.c_double),
        ("init_bb0", C.c_uint64), ("init_bb1", C.c_uint64),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        _lib.c4o_abi_version.restype = C.c_int
    return _lib


def set_threads(n: int) -> int:
    """Use `n` OpenMP threads from now on (overrides OMP_NUM_THREADS); returns the number in effect."""
    l = lib()
    l.c4o_set_threads.restype = C.c_int
    l.c4o_set_threads.argtypes = [C.c_int]
    return int(l.c4o_set_threads(int(n)))


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _wrap_cb(py_eval):
    """py_eval(bb0[n], bb1[n], player[n], legal[n]) -> (priors[n,7] f32, values[n,2] f32)"""
    if py_eval is None:
        return C.cast(None, EVAL_CB), None

    def tramp(user, n, b0, b1, pl, lg, pri, val):
        bb0 = np.ctypeslib.as_array(b0, (n,)).copy()
        bb1 = np.ctypeslib.as_array(b1, (n,)).copy()
        player = np.ctypeslib.as_array(pl, (n,)).copy()
        legal = np.ctypeslib.as_array(lg, (n,)).copy()
        p, v = py_eval(bb0, bb1, player, legal)
        np.ctypeslib.as_array(pri, (n, 7))[:] = np.asarray(p, dtype=np.float32).reshape(n, 7)
        np.ctypeslib.as_array(val, (n, 2))[:] = np.asarray(v, dtype=np.float32).reshape(n, 2)

    cb = EVAL_CB(tramp)
    return cb, tramp


def env_step(bb0, bb1, player, col):
    n = len(bb0)
    bb0 = np.ascontiguousarray(bb0, np.uint64); bb1 = np.ascontiguousarray(bb1, np.uint64)
    player = np.ascontiguousarray(player, np.uint8); col = np.ascontiguousarray(col, np.uint8)
    o0 = np.empty(n, np.uint64); o1 = np.empty(n, np.uint64); op = np.empty(n, np.uint8)
    lg = np.empty(n, np.uint8); en = np.empty(n, np.uint8); rw = np.empty((n, 2), np.int8); st = np.empty(n, np.uint8)
    lib().c4o_env_step(C.c_int(n), _p(bb0, C.c_uint64), _p(bb1, C.c_uint64), _p(player, C.c_uint8), _p(col, C.c_uint8),
                       _p(o0, C.c_uint64), _p(o1, C.c_uint64), _p(op, C.c_uint8), _p(lg, C.c_uint8), _p(en, C.c_uint8),
                       _p(rw, C.c_int8), _p(st, C.c_uint8))
    return dict(bb0=o0, bb1=o1, player=op, legal=lg, ended=en, reward=rw, status=st)


def state_info(bb0, bb1, player):
    n = len(bb0)
    bb0 = np.ascontiguousarray(bb0, np.uint64); bb1 = np.ascontiguousarray(bb1, np.uint64)
    player = np.ascontiguousarray(player, np.uint8)
    lg = np.empty(n, np.uint8); en = np.empty(n, np.uint8); rw = np.empty((n, 2), np.int8)
    lib().c4o_state_info(C.c_int(n), _p(bb0, C.c_uint64), _p(bb1, C.c_uint64), _p(player, C.c_uint8),
                         _p(lg, C.c_uint8), _p(en, C.c_uint8), _p(rw, C.c_int8))
    return dict(legal=lg, ended=en, reward=rw)


def search(bb0, bb1, player, S, c_puct=1.0, eval_kind=1, py_eval=None):
    E = len(bb0)
    bb0 = np.ascontiguousarray(bb0, np.uint64); bb1 = np.ascontiguousarray(bb1, np.uint64)
    player = np.ascontiguousarray(player, np.uint8)
    cN = np.zeros((E, 7), np.int32); cW = np.zeros((E, 7), np.float64); cP = np.zeros((E, 7), np.float32)
    rW = np.zeros(E, np.float64); rN = np.zeros(E, np.int32); lg = np.zeros(E, np.uint8)
    nev = C.c_int64(0)
    cb, keep = _wrap_cb(py_eval)
    err = lib().c4o_search(C.c_int(E), _p(bb0, C.c_uint64), _p(bb1, C.c_uint64), _p(player, C.c_uint8), C.c_int(S),
                           C.c_double(c_puct), C.c_int(eval_kind), cb, None, _p(cN, C.c_int32), _p(cW, C.c_double),
                           _p(cP, C.c_float), _p(rW, C.c_double), _p(rN, C.c_int32), _p(lg, C.c_uint8), C.byref(nev))
    if err:
        raise RuntimeError(f"c4o_search error {err}")
    return dict(child_N=cN, child_W=cW, child_P=cP, root_W=rW, root_N=rN, legal=lg, n_evals=nev.value)


@dataclass
class SelfPlayResult:
    ep_slot: np.ndarray
    ep_len: np.ndarray
    ep_step: np.ndarray
    ep_outcome: np.ndarray  # [n_ep, 2] int8
    s_bb0: np.ndarray
    s_bb1: np.ndarray
    s_player: np.ndarray
    s_counts: np.ndarray  # [n_samples, 7] int32 root-child visit counts (policy = counts / (S-1))
    n_steps: int
    n_uniforms_used: int
    n_sims: int
    n_evals: int


def selfplay(E, S, uniforms, quota=None, c_puct=1.0, eval_kind=1, init=(0, 0, 0), py_eval=None) -> SelfPlayResult:
    quota = E if quota is None else quota
    uniforms = np.ascontiguousarray(uniforms, np.float64).reshape(-1, E)
    max_steps = uniforms.shape[0]
    n_ep_cap = int(min(quota, E * max_steps // 7 + E))  # a game lasts at least 7 plies
    max_samples = n_ep_cap * 42
    quota = int(min(quota, 2**31 - 1))
    ep_slot = np.zeros(n_ep_cap, np.int32); ep_len = np.zeros(n_ep_cap, np.int32); ep_step = np.zeros(n_ep_cap, np.int32)
    ep_out = np.zeros((n_ep_cap, 2), np.int8)
    cfg = _Cfg(E, S, eval_kind, init[2], quota, max_steps, c_puct, init[0], init[1])
    s0 = np.zeros(max_samples, np.uint64); s1 = np.zeros(max_samples, np.uint64); sp = np.zeros(max_samples, np.uint8)
    sc = np.zeros((max_samples, 7), np.int32)
    ne, ns, nst, nu, nsim, nev = (C.c_int64(0) for _ in range(6))
    cb, keep = _wrap_cb(py_eval)
    err = lib().c4o_selfplay(C.byref(cfg), _p(uniforms, C.c_double), cb, None, _p(ep_slot, C.c_int32), _p(ep_len, C.c_int32),
                             _p(ep_step, C.c_int32), _p(ep_out, C.c_int8), _p(s0, C.c_uint64), _p(s1, C.c_uint64),
                             _p(sp, C.c_uint8), _p(sc, C.c_int32), C.c_int64(max_samples), C.byref(ne), C.byref(ns),
                             C.byref(nst), C.byref(nu), C.byref(nsim), C.byref(nev))
    if err:
        raise RuntimeError(f"c4o_selfplay error {err}")
    n_ep, n_s = ne.value, ns.value
    return SelfPlayResult(ep_slot[:n_ep], ep_len[:n_ep], ep_step[:n_ep], ep_out[:n_ep], s0[:n_s], s1[:n_s], sp[:n_s],
                          sc[:n_s], nst.value, nu.value, nsim.value, nev.value)
