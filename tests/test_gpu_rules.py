"""GPU parity: Connect4 bitboard kernels vs the reference-derived goldens and the C oracle (bit-exact)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from alphazero_implementation_b200 import Engine

    e = Engine(num_games=8, num_simulations=4)
    yield e
    e.close()


def _np(d):
    return {k: v.cpu().numpy() for k, v in d.items()}


def test_env_step_golden_playouts(eng, rules_goldens):
    b0, b1, pl, col, exp = [], [], [], [], []
    for plies in rules_goldens["games"]:
        p0, p1, pp = 0, 0, 0
        for p in plies:
            b0.append(p0); b1.append(p1); pl.append(pp); col.append(p["col"]); exp.append(p)
            p0, p1, pp = p["bb0"], p["bb1"], p["player"]
    r = _np(eng.env_step(np.array(b0, np.uint64), np.array(b1, np.uint64), np.array(pl, np.uint8), np.array(col, np.uint8)))
    assert (r["status"] == 0).all()
    assert r["bb0"].view(np.uint64).tolist() == [p["bb0"] for p in exp]
    assert r["bb1"].view(np.uint64).tolist() == [p["bb1"] for p in exp]
    assert r["player"].tolist() == [p["player"] for p in exp]
    assert r["ended"].tolist() == [p["ended"] for p in exp]
    assert r["legal"].tolist() == [0 if p["ended"] else p["legal"] for p in exp]
    assert r["reward"].tolist() == [p["reward"] for p in exp]


def test_env_step_random_vs_oracle(eng, oracle):
    """200k random positions reached by random play (incl. finished games and full columns) + random columns."""
    rng = np.random.RandomState(7)
    n = 200_000
    b0 = np.zeros(n, np.uint64); b1 = np.zeros(n, np.uint64); pl = np.zeros(n, np.uint8)
    depth = rng.randint(0, 43, n)
    for d in range(42):  # advance the positions that still want moves, on the oracle (CPU)
        act = np.nonzero(depth > d)[0]
        if len(act) == 0:
            break
        r = oracle.env_step(b0[act], b1[act], pl[act], rng.randint(0, 7, len(act)).astype(np.uint8))
        b0[act], b1[act], pl[act] = r["bb0"], r["bb1"], r["player"]  # illegal moves leave the state unchanged
    col = rng.randint(0, 9, n).astype(np.uint8)  # 7, 8 = out of range
    ref = oracle.env_step(b0, b1, pl, col)
    got = _np(eng.env_step(b0, b1, pl, col))
    for k in ("status", "player", "legal", "ended"):
        assert (got[k] == ref[k]).all(), k
    assert (got["bb0"].view(np.uint64) == ref["bb0"]).all() and (got["bb1"].view(np.uint64) == ref["bb1"]).all()
    assert (got["reward"] == ref["reward"]).all()
    assert ref["ended"].sum() > 1000 and (ref["status"] == 1).sum() > 1000
    info_ref = oracle.state_info(b0, b1, pl)
    info = _np(eng.state_info(b0, b1))
    for k in ("legal", "ended", "reward"):
        assert (info[k] == info_ref[k]).all(), k


def test_env_step_empty_batch(eng):
    r = eng.env_step(np.zeros(0, np.uint64), np.zeros(0, np.uint64), np.zeros(0, np.uint8), np.zeros(0, np.uint8))
    assert r["bb0"].numel() == 0


def test_plane_encoders_match_reference_definitions(eng, nets_goldens, rules_goldens):
    from alphazero_implementation_b200.engine import (LAYOUT_GRID_F32, LAYOUT_PLANES_BF16, LAYOUT_PLANES_BF16_NHWC,
                                                      LAYOUT_PLANES_F32)
    from alphazero_implementation_b200.game import bitboards_to_grid

    sts = nets_goldens["states"] + [dict(bb0=p["bb0"], bb1=p["bb1"], player=p["player"]) for g in rules_goldens["games"][:40] for p in g]
    b0 = np.array([s["bb0"] for s in sts], np.uint64); b1 = np.array([s["bb1"] for s in sts], np.uint64)
    pl = np.array([s["player"] for s in sts], np.uint8)
    grids = np.stack([bitboards_to_grid(int(a), int(b)) for a, b in zip(b0, b1)]).astype(np.float32)
    g = eng.encode_states(b0, b1, pl, LAYOUT_GRID_F32).cpu().numpy()
    assert (g == grids).all()  # BasicNN._states_to_tensor (basic.py:41-47)
    planes = np.stack([grids == -1, grids == pl[:, None, None], grids == (1 - pl)[:, None, None]], axis=1).astype(np.float32)
    p32 = eng.encode_states(b0, b1, pl, LAYOUT_PLANES_F32).cpu().numpy()
    assert (p32 == planes).all()  # CNNModel._states_to_tensor (cnn.py:77-100)
    p16 = eng.encode_states(b0, b1, pl, LAYOUT_PLANES_BF16).float().cpu().numpy()
    assert (p16 == planes).all()
    nhwc = eng.encode_states(b0, b1, pl, LAYOUT_PLANES_BF16_NHWC).float().cpu().numpy()
    assert (nhwc[..., :3] == planes.transpose(0, 2, 3, 1)).all() and (nhwc[..., 3:] == 0).all()
    # golden: first BasicNN input row and the input checksum recorded from the reference
    n0 = len(nets_goldens["states"])
    assert g[0].flatten().tolist() == nets_goldens["BasicNN"]["input_first"]
    assert float(g[:n0].sum()) == nets_goldens["BasicNN"]["input_sum"]
    assert float(p32[:n0].sum()) == nets_goldens["CNNModel"]["input_sum"]


def test_masked_softmax_matches_torch(eng):
    torch.manual_seed(0)
    logits = torch.randn(5000, 7) * 3
    legal = torch.randint(1, 128, (5000,), dtype=torch.uint8)
    got = eng.masked_softmax(logits, legal).cpu()
    for i in range(0, 5000, 97):
        cols = [c for c in range(7) if (int(legal[i]) >> c) & 1]
        ref = torch.softmax(logits[i, cols], dim=0)  # model.py:29-35
        assert torch.allclose(got[i, cols], ref, atol=1e-6, rtol=1e-5)
        assert got[i].sum().item() == pytest.approx(1.0, abs=1e-5)
        assert all(got[i, c] == 0 for c in range(7) if c not in cols)
    # 16-byte-aligned batches take the shared-memory-staged kernel for the full 256-row blocks; a misaligned view of the same
    # rows takes the thread-per-row kernel.  Same arithmetic: bit-identical.
    dl, dg = logits.cuda(), legal.cuda()
    assert dl[1:].data_ptr() % 16 != 0
    assert torch.equal(eng.masked_softmax(dl[1:], dg[1:]).cpu(), got[1:])


def test_env_step_vector_and_scalar_paths_agree(eng, oracle):
    """16-byte-aligned batches take the 4-positions-per-thread kernel, anything else the scalar one (plus a scalar tail)."""
    rng = np.random.RandomState(3)
    n = 4099
    b0 = np.zeros(n, np.uint64); b1 = np.zeros(n, np.uint64); pl = np.zeros(n, np.uint8)
    for d in range(14):
        r = oracle.env_step(b0, b1, pl, rng.randint(0, 7, n).astype(np.uint8))
        b0, b1, pl = r["bb0"], r["bb1"], r["player"]
    col = rng.randint(0, 8, n).astype(np.uint8)
    ref = oracle.env_step(b0, b1, pl, col)
    tb0 = torch.from_numpy(b0.view(np.int64)).cuda(); tb1 = torch.from_numpy(b1.view(np.int64)).cuda()
    tpl = torch.from_numpy(pl).cuda(); tcol = torch.from_numpy(col).cuda()
    for off in (0, 1, 2, 3):  # off != 0: 8-byte aligned bitboards, odd byte arrays -> scalar kernel
        got = _np(eng.env_step(tb0[off:], tb1[off:], tpl[off:], tcol[off:]))
        for k in ("status", "player", "legal", "ended"):
            assert (got[k] == ref[k][off:]).all(), (off, k)
        assert (got["bb0"].view(np.uint64) == ref["bb0"][off:]).all() and (got["reward"] == ref["reward"][off:]).all()
        info = _np(eng.state_info(tb0[off:], tb1[off:]))
        iref = oracle.state_info(b0[off:], b1[off:], pl[off:])
        for k in ("legal", "ended", "reward"):
            assert (info[k] == iref[k]).all(), (off, k)


@pytest.mark.parametrize("n", [1, 63, 64, 65, 129, 1000])
def test_plane_encoder_tile_edges(eng, oracle, n):
    from alphazero_implementation_b200.engine import LAYOUT_GRID_F32, LAYOUT_PLANES_BF16, LAYOUT_PLANES_BF16_NHWC, LAYOUT_PLANES_F32
    from alphazero_implementation_b200.game import bitboards_to_grid

    rng = np.random.RandomState(n)
    b0 = np.zeros(n, np.uint64); b1 = np.zeros(n, np.uint64); pl = np.zeros(n, np.uint8)
    for d in range(9):
        r = oracle.env_step(b0, b1, pl, rng.randint(0, 7, n).astype(np.uint8))
        b0, b1, pl = r["bb0"], r["bb1"], r["player"]
    grids = np.stack([bitboards_to_grid(int(a), int(b)) for a, b in zip(b0, b1)]).astype(np.float32)
    planes = np.stack([grids == -1, grids == pl[:, None, None], grids == (1 - pl)[:, None, None]], axis=1).astype(np.float32)
    assert (eng.encode_states(b0, b1, pl, LAYOUT_GRID_F32).cpu().numpy() == grids).all()
    assert (eng.encode_states(b0, b1, pl, LAYOUT_PLANES_F32).cpu().numpy() == planes).all()
    assert (eng.encode_states(b0, b1, pl, LAYOUT_PLANES_BF16).float().cpu().numpy() == planes).all()
    nhwc = eng.encode_states(b0, b1, pl, LAYOUT_PLANES_BF16_NHWC).float().cpu().numpy()
    assert (nhwc[..., :3] == planes.transpose(0, 2, 3, 1)).all() and (nhwc[..., 3:] == 0).all()
