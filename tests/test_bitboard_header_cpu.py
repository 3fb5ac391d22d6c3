"""CPU check of the product's rules primitives: `csrc/c4_bitboard.cuh` is host/device code (the tree, env and encoder
kernels all inline it), so it is compiled here with g++ behind a tiny extern "C" shim and compared with the oracle on
reachable positions (legal masks, drops, terminal / reward) and on arbitrary stone sets (branch-free line test ≡ branching one).
No GPU needed."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "alphazero-implementation_b200", "csrc", "c4_bitboard.cuh")

SHIM = r'''
#include "%s"
extern "C" {
void c4h_info(int n, const uint64_t *b0, const uint64_t *b1, uint8_t *legal, uint8_t *ended, int8_t *reward0) {
    for (int i = 0; i < n; ++i) {
        c4::Terminal t = c4::terminal_of(b0[i], b1[i]);
        ended[i] = t.ended;
        reward0[i] = t.reward0;
        legal[i] = t.ended ? 0 : (uint8_t)c4::legal_mask(b0[i] | b1[i]);
    }
}
void c4h_drop(int n, const uint64_t *b0, const uint64_t *b1, const uint8_t *pl, const uint8_t *col, uint64_t *o0, uint64_t *o1) {
    for (int i = 0; i < n; ++i) {
        const uint64_t bit = c4::drop_bit(b0[i] | b1[i], col[i]);
        o0[i] = b0[i] | (pl[i] == 0 ? bit : 0);
        o1[i] = b1[i] | (pl[i] == 0 ? 0 : bit);
    }
}
void c4h_lines(int n, const uint64_t *b, uint8_t *branching, uint8_t *branch_free, uint8_t *full) {
    for (int i = 0; i < n; ++i) {
        branching[i] = c4::has4(b[i]);
        branch_free[i] = c4::has4_nb(b[i]);
        full[i] = c4::is_full(b[i]);
    }
}
int c4h_nth(uint32_t legal, int idx) { return c4::nth_legal_column(legal, idx); }
}
'''


@pytest.fixture(scope="module")
def hdr(tmp_path_factory):
    d = tmp_path_factory.mktemp("c4h")
    src = d / "shim.cpp"
    src.write_text(SHIM % HEADER)
    so = d / "libc4h.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", str(so), str(src)], check=True)
    return C.CDLL(str(so))


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def test_header_rules_match_the_oracle_on_reachable_positions(hdr, oracle):
    rng = np.random.RandomState(11)
    n = 20000
    b0 = np.zeros(n, np.uint64); b1 = np.zeros(n, np.uint64); pl = np.zeros(n, np.uint8)
    seen_end = 0
    for ply in range(42):
        info = oracle.state_info(b0, b1, pl)
        lg = np.empty(n, np.uint8); en = np.empty(n, np.uint8); r0 = np.empty(n, np.int8)
        hdr.c4h_info(n, _p(b0, C.c_uint64), _p(b1, C.c_uint64), _p(lg, C.c_uint8), _p(en, C.c_uint8), _p(r0, C.c_int8))
        assert (lg == info["legal"]).all() and (en == info["ended"]).all() and (r0 == info["reward"][:, 0]).all(), ply
        seen_end += int(en.sum())
        # a random LEGAL column where the game goes on (finished games keep their position: the oracle rejects the move)
        col = np.array([rng.choice([c for c in range(7) if (m >> c) & 1]) if m else 0 for m in lg], np.uint8)
        ref = oracle.env_step(b0, b1, pl, col)
        o0 = np.empty(n, np.uint64); o1 = np.empty(n, np.uint64)
        hdr.c4h_drop(n, _p(b0, C.c_uint64), _p(b1, C.c_uint64), _p(pl, C.c_uint8), _p(col, C.c_uint8), _p(o0, C.c_uint64), _p(o1, C.c_uint64))
        live = en == 0
        assert (o0[live] == ref["bb0"][live]).all() and (o1[live] == ref["bb1"][live]).all(), ply
        assert (ref["status"][live] == 0).all() and (ref["status"][~live] == 1).all()
        b0, b1, pl = ref["bb0"], ref["bb1"], ref["player"]
    assert seen_end > n  # most games ended well before ply 42 and were re-checked every ply after that


def test_branch_free_line_test_equals_the_branching_one(hdr):
    rng = np.random.RandomState(5)
    board = np.uint64(0xFDFBF7EFDFBF)  # the 42 playable cells (guard bit of every column clear)
    n = 200000
    dens = rng.random_sample(n)
    bits = (rng.random_sample((n, 49)) < dens[:, None]).astype(np.uint64)
    b = (bits << np.arange(49, dtype=np.uint64)).sum(axis=1).astype(np.uint64) & board
    a = np.empty(n, np.uint8); nb = np.empty(n, np.uint8); full = np.empty(n, np.uint8)
    hdr.c4h_lines(n, _p(b, C.c_uint64), _p(a, C.c_uint8), _p(nb, C.c_uint8), _p(full, C.c_uint8))
    assert (a == nb).all() and 0 < int(a.sum()) < n
    # straightforward cell-by-cell line test on a sample
    for i in rng.choice(n, 300, replace=False):
        g = [[(int(b[i]) >> (7 * c + r)) & 1 for r in range(6)] for c in range(7)]
        want = any(all(0 <= c + k * dc < 7 and 0 <= r + k * dr < 6 and g[c + k * dc][r + k * dr] for k in range(4))
                   for c in range(7) for r in range(6) for dc, dr in ((0, 1), (1, 0), (1, 1), (1, -1)))
        assert bool(a[i]) == want
    top = np.uint64(0x810204081020)
    assert (full == ((b & top) == top)).all()


def test_nth_legal_column(hdr):
    for legal in range(1, 128):
        cols = [c for c in range(7) if (legal >> c) & 1]
        assert [hdr.c4h_nth(legal, i) for i in range(len(cols))] == cols
