"""GPU parity: tree arena kernels vs goldens produced by the reference's own AlphaZeroSearch, and vs the
C oracle at BASELINE config-2 size.  Everything here is bit-exact (visit counts, fp64 value sums, fp32 priors)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from alphazero_implementation_b200.engine import POLICY_PRIORS, Engine  # noqa: E402


def _np(d):
    return {k: v.cpu().numpy() for k, v in d.items()}


@pytest.mark.parametrize("lanes", [8, 32])
def test_search_goldens_fused(search_goldens, lanes):
    groups = {}
    for c in search_goldens:
        groups.setdefault((c["S"], c["c_puct"], c["eval_kind"]), []).append(c)
    for (S, cp, kind), cases in groups.items():
        eng = Engine(num_games=len(cases), num_simulations=S, c_puct=cp, lanes_per_tree=lanes)
        eng.set_roots(np.array([c["bb0"] for c in cases], np.uint64), np.array([c["bb1"] for c in cases], np.uint64),
                      np.array([c["player"] for c in cases], np.uint8))
        eng.run_simulations(S, kind)
        st = _np(eng.root_stats())
        for i, c in enumerate(cases):
            assert st["child_N"][i].tolist() == c["child_N"], c["name"]
            assert st["child_W"][i].tolist() == c["child_W"], c["name"]
            assert [float(x) for x in st["child_P"][i]] == c["child_P"], c["name"]
            assert st["root_N"][i] == c["root_N"] and st["root_W"][i] == c["root_W"], c["name"]
            assert st["legal"][i] == c["legal"] and st["err"][i] == 0
        s = eng.stats()
        assert s["simulations"] == S * len(cases) and s["backup_nodes"] == s["levels"] + s["simulations"]
        eng.close()


def _random_roots(oracle, n, seed, max_depth=30):
    rng = np.random.RandomState(seed)
    b0 = np.zeros(n, np.uint64); b1 = np.zeros(n, np.uint64); pl = np.zeros(n, np.uint8)
    depth = rng.randint(0, max_depth, n)
    for d in range(max_depth):
        act = np.nonzero(depth > d)[0]
        if len(act) == 0:
            break
        r = oracle.env_step(b0[act], b1[act], pl[act], rng.randint(0, 7, len(act)).astype(np.uint8))
        ok = r["ended"] == 0  # never step into a finished game: roots must be live positions
        idx = act[ok]
        b0[idx], b1[idx], pl[idx] = r["bb0"][ok], r["bb1"][ok], r["player"][ok]
    return b0, b1, pl


@pytest.mark.parametrize("lanes,kind,S,n", [(8, 2, 16, 129), (8, 1, 40, 1), (16, 2, 33, 3), (32, 1, 12, 5), (16, 2, 200, 2048), (8, 1, 200, 4096), (32, 2, 200, 4096), (8, 2, 800, 512), (32, 1, 50, 1000), (8, 2, 37, 777)])
def test_search_vs_oracle_config2_size(oracle, lanes, kind, S, n):
    """4096 trees x 200 sims (BASELINE config 2) from random mid-game roots, compared with the C oracle."""
    b0, b1, pl = _random_roots(oracle, n, seed=S + n)
    ref = oracle.search(b0, b1, pl, S, c_puct=1.0, eval_kind=kind)
    eng = Engine(num_games=n, num_simulations=S, lanes_per_tree=lanes)
    eng.set_roots(b0, b1, pl)
    eng.run_simulations(S, kind)
    st = _np(eng.root_stats())
    assert (st["child_N"] == ref["child_N"]).all()
    assert (st["child_W"] == ref["child_W"]).all()
    assert (st["child_P"] == ref["child_P"]).all()
    assert (st["root_N"] == ref["root_N"]).all() and (st["root_W"] == ref["root_W"]).all()
    assert (st["legal"] == ref["legal"]).all()
    assert eng.stats()["evaluations"] == ref["n_evals"]
    assert (st["child_N"].sum(1) == S - 1).all() and (st["root_N"] == S).all()  # size-independent invariant (App. A.4)
    eng.close()


@pytest.mark.parametrize("lanes", [8, 16, 32])
def test_split_path_equals_fused(oracle, lanes):
    """select -> (evaluator) -> expand/backup with exact priors reproduces the fused kernel's trees."""
    from oracle import evaluators as ev

    n, S, kind = 96, 60, 2
    b0, b1, pl = _random_roots(oracle, n, seed=5)
    fused = Engine(num_games=n, num_simulations=S, lanes_per_tree=lanes)
    fused.set_roots(b0, b1, pl); fused.run_simulations(S, kind)
    a = _np(fused.root_stats())
    split = Engine(num_games=n, num_simulations=S, lanes_per_tree=lanes)
    split.set_roots(b0, b1, pl)
    for _ in range(S):
        split.select_leaves()
        info = _np(split.leaf_info())
        pri = np.zeros((n, 7), np.float32); val = np.zeros((n, 2), np.float32)
        for i in np.nonzero(info["status"] == 0)[0]:
            cols = [c for c in range(7) if (info["legal"][i] >> c) & 1]
            p, v = ev.evaluate(kind, int(info["bb0"][i]), int(info["bb1"][i]), int(info["player"][i]), cols)
            for c, x in p.items():
                pri[i, c] = x
            val[i] = v
        split.expand_backup(torch.from_numpy(pri).cuda(), torch.from_numpy(val).cuda(), POLICY_PRIORS)
    b = _np(split.root_stats())
    for k in ("child_N", "child_W", "child_P", "root_N", "root_W"):
        assert (a[k] == b[k]).all(), k
    ta, tb = fused.export_tree(3), split.export_tree(3)
    assert ta["used"] == tb["used"]
    for k in ("W", "N", "P", "first_child"):
        assert (ta[k] == tb[k]).all()
    sa, sb = fused.stats(), split.stats()
    for k in ("simulations", "evaluations", "levels", "children_created"):
        assert sa[k] == sb[k], k
    fused.close(); split.close()


def test_terminal_root_is_flagged(oracle):
    # player 0 has four in the bottom row
    b0 = np.array([(1 << 0) | (1 << 7) | (1 << 14) | (1 << 21), 0], np.uint64)
    b1 = np.array([(1 << 1) | (1 << 8) | (1 << 15), 0], np.uint64)
    eng = Engine(num_games=2, num_simulations=10)
    eng.set_roots(b0, b1, np.array([1, 0], np.uint8))
    eng.run_simulations(10, 1)
    st = _np(eng.root_stats())
    assert st["err"].tolist() == [3, 0] and st["root_N"].tolist() == [0, 10]
    eng.close()


def test_gather_leaves_layouts(oracle):
    from alphazero_implementation_b200.engine import LAYOUT_GRID_F32, LAYOUT_PLANES_BF16, LAYOUT_PLANES_F32
    from alphazero_implementation_b200.game import bitboards_to_grid

    n = 333
    b0, b1, pl = _random_roots(oracle, n, seed=11)
    eng = Engine(num_games=n, num_simulations=8)
    eng.set_roots(b0, b1, pl)
    eng.run_simulations(5, 2)
    eng.select_leaves()
    info = _np(eng.leaf_info())
    grids = np.stack([bitboards_to_grid(int(a), int(b)) for a, b in zip(info["bb0"], info["bb1"])]).astype(np.float32)
    live = info["status"] == 0
    lp = info["player"].astype(np.int64)
    planes = np.stack([grids == -1, grids == lp[:, None, None], grids == (1 - lp)[:, None, None]], axis=1).astype(np.float32)
    grids[~live] = 0; planes[~live] = 0
    assert (eng.gather_leaves(LAYOUT_GRID_F32).cpu().numpy() == grids).all()
    assert (eng.gather_leaves(LAYOUT_PLANES_F32).cpu().numpy() == planes).all()
    assert (eng.gather_leaves(LAYOUT_PLANES_BF16).float().cpu().numpy() == planes).all()
    assert live.sum() > 0
    eng.close()


@pytest.mark.parametrize("S", [200, 800, 20000])
def test_table_division_is_correctly_rounded(S):
    """PUCT divides by small integers through a reciprocal table + two FMA corrections; the result must equal
    the IEEE quotient (what CPython computes) for every (x, d): 2^28 pseudo-random pairs per table size."""
    eng = Engine(num_games=8, num_simulations=S)
    assert eng.selftest_division(1 << 28, seed=S) == 0
    eng.close()


def test_simulation_budget_is_enforced():
    eng = Engine(num_games=4, num_simulations=50)
    eng.run_simulations(30, 1)
    eng.run_simulations(20, 1)  # continuing the same roots is fine up to the arena size
    with pytest.raises(RuntimeError):
        eng.run_simulations(1, 1)
    assert eng.root_stats()["root_N"].cpu().tolist() == [50] * 4
    eng.reset_games()
    eng.run_simulations(50, 2)
    eng.close()


@pytest.mark.parametrize("E,S", [(16384, 800), (65536, 800)])
def test_full_size_configs_by_invariants(oracle, E, S):
    """BASELINE configs 3 / 4 sizes (16384 and 65536 trees x 800 sims): too large for the CPU oracle in a test, so
    checked through size-independent properties — every root has N = S and child visits summing to S - 1 (App. A.4),
    visits only on legal columns, identical roots give identical trees, and a random sample of trees is bit-exact
    against the oracle."""
    rng = np.random.RandomState(E)
    n_distinct = 64
    b0d, b1d, pld = _random_roots(oracle, n_distinct, seed=E, max_depth=20)
    pick = rng.randint(0, n_distinct, E)
    b0, b1, pl = b0d[pick], b1d[pick], pld[pick]
    eng = Engine(num_games=E, num_simulations=S)
    eng.set_roots(b0, b1, pl)
    eng.run_simulations(S, 2)
    st = _np(eng.root_stats())
    assert (st["root_N"] == S).all() and (st["child_N"].sum(1) == S - 1).all() and (st["err"] == 0).all()
    legal_bits = ((st["legal"][:, None] >> np.arange(7)[None, :]) & 1).astype(bool)
    assert (st["child_N"][~legal_bits] == 0).all()
    # identical roots -> identical statistics, wherever the slot sits in the arena
    first = {}
    for i, p in enumerate(pick[:4096]):
        j = first.setdefault(int(p), i)
        assert (st["child_N"][i] == st["child_N"][j]).all() and (st["child_W"][i] == st["child_W"][j]).all()
    ref = oracle.search(b0d, b1d, pld, S, eval_kind=2)
    for p in range(n_distinct):
        i = first.get(p)
        if i is None:
            continue
        assert (st["child_N"][i] == ref["child_N"][p]).all() and (st["child_W"][i] == ref["child_W"][p]).all()
    s = eng.stats()
    assert s["simulations"] == E * S and s["backup_nodes"] == s["levels"] + s["simulations"]
    eng.close()
