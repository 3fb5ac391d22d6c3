"""TEST INFRASTRUCTURE — generate tests/golden/*.json by running the REFERENCE ITSELF.

Run in the build container only (needs /root/reference):
    python -m oracle.gen_golden

The reference's unmodified `AlphaZeroSearch` / `Node` / `EpisodeGenerator`
(`core/search/mcts/search.py`, `node.py`, `core/training/episode_generator.py`)
are imported from /root/reference/src with the shims in oracle/shims/ and driven
with the deterministic evaluators of oracle/evaluators.py.  The emitted vectors
pin (a) the C restatement `c4_oracle.c` and (b) the CUDA engine.

Files
  search_goldens.json    root statistics of `AlphaZeroSearch.run` on the empty board and on
                         every fixture position of notebooks/policy_comparison.ipynb#cell6/#cell11.
  selfplay_goldens.json  full `generate_episodes` transcripts (per-move visit counts, states,
                         outcomes, yield order, number of RNG draws) under np.random.seed(seed).
  rules_goldens.json     random playouts through the game-rules stand-in, cross-checked move by move
                         against the in-tree rules `src/alphazero_simple/connect4_game.py:28-98`.
  training_goldens.json  losses of four `training_step` + Adam steps of BasicNN / CNNModel on one fixed batch.
  basicnn_goldens.json   `BasicNN.predict` / `CNNModel.predict` outputs (fp32, CPU) on fixture states
                         with seeded random weights, plus one C1 run (E=1, S=100, BasicNN).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.evaluators import HASH, UNIFORM, DeterministicEvaluator, grid_to_bitboards  # noqa: E402
from oracle.ref_loader import REFERENCE_SRC, load_reference  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
E_ = -1

# notebooks/policy_comparison.ipynb#cell6 (row 0 = bottom, -1 empty)
FINAL_SITUATIONS = [
    dict(grid=[[0, 0, 0, E_, E_, 1, E_], [E_, E_, E_, E_, E_, 1, E_], [E_, E_, E_, E_, E_, 1, E_]], player=0, expected_move=3),
    dict(grid=[[1, 1, 1, E_, E_, 0, E_], [E_, E_, E_, E_, E_, 0, E_], [E_, E_, E_, E_, E_, 0, E_]], player=0, expected_move=5),
    dict(grid=[[E_, 0, 0, 0, E_, E_, E_], [E_, 1, 1, 1, E_, E_, E_]], player=0, expected_move=0),
    dict(grid=[[0, 0, 0, E_, E_, 1, E_], [E_, 0, E_, E_, E_, 1, E_], [E_, E_, E_, E_, E_, 1, E_]], player=1, expected_move=5),
    dict(grid=[[0, 1, 0, 1, E_, E_, E_], [1, 0, 1, 0, E_, E_, E_], [0, E_, E_, E_, E_, E_, E_], [1, E_, E_, E_, E_, E_, E_]], player=1, expected_move=1),
    dict(grid=[[0, 1, 0, 1, E_, 0, 1], [0, 0, 0, E_, E_, 1, 0], [1, 1, E_, E_, E_, 1, E_]], player=0, expected_move=3),
    dict(grid=[[0, 1, 0, 1, 1, E_, E_], [1, 0, 0, 1, E_, E_, E_], [0, 0, 1, 0, E_, E_, E_]], player=1, expected_move=1),
]
# notebooks/policy_comparison.ipynb#cell11
START_SITUATIONS = [
    dict(grid=[], player=0),
    dict(grid=[[0, E_, E_, E_, E_, E_, E_]], player=1),
    dict(grid=[[0, E_, E_, E_, 1, E_, E_]], player=0),
    dict(grid=[[0, E_, E_, E_, 1, E_, E_], [E_, E_, E_, E_, 0, E_, E_]], player=1),
    dict(grid=[[0, E_, E_, E_, 1, E_, 1], [E_, E_, E_, E_, 0, E_, E_]], player=0),
    dict(grid=[[0, E_, E_, E_, 1, E_, 1], [E_, E_, E_, E_, 0, E_, E_], [E_, E_, E_, E_, 0, E_, E_]], player=1),
    dict(grid=[[0, E_, E_, E_, 1, 1, 1], [E_, E_, E_, E_, 0, E_, E_], [E_, E_, E_, E_, 0, E_, E_]], player=0, expected_move=3),
    dict(grid=[[0, 0, E_, E_, 1, 1, 1], [E_, E_, E_, E_, 0, E_, E_], [E_, E_, E_, E_, 0, E_, E_]], player=1, expected_move=3),
    dict(grid=[[0, E_, E_, 0, 1, 1, 1], [E_, E_, E_, E_, 0, E_, E_], [E_, E_, E_, E_, 0, E_, E_]], player=1),
    dict(grid=[[0, E_, E_, 0, 1, 1, 1], [E_, E_, E_, E_, 0, 1, E_], [E_, E_, E_, E_, 0, E_, E_]], player=0),
    dict(grid=[[0, E_, 0, 0, 1, 1, 1], [E_, E_, E_, E_, 0, 1, E_], [E_, E_, E_, E_, 0, E_, E_]], player=1),
    dict(grid=[[0, E_, 0, 0, 1, 1, 1], [E_, E_, E_, E_, 0, 1, E_], [E_, E_, E_, E_, 0, 1, E_]], player=0),
]


def full_grid(rows):
    g = [list(r) for r in rows]
    while len(g) < 6:
        g.append([E_] * 7)
    return g


def state_json(sit):
    return {"config": {"count": 4, "height": 6, "width": 7}, "grid": full_grid(sit["grid"]), "player": sit["player"]}


def legal_mask(state):
    m = 0
    for a in state.actions:
        m |= 1 << a.column
    return m


def gen_search(R):
    cases = []
    positions = [("empty", dict(grid=[], player=0))]
    positions += [(f"final{i}", s) for i, s in enumerate(FINAL_SITUATIONS)]
    positions += [(f"start{i}", s) for i, s in enumerate(START_SITUATIONS)]
    for name, sit in positions:
        sims = (100, 200, 800) if name == "empty" else (100, 300)
        for kind in (UNIFORM, HASH):
            for S in sims:
                for c_puct in ((1.0, 2.5) if name in ("empty", "final5") and S == 100 else (1.0,)):
                    state = R.State.from_json(state_json(sit))
                    root = R.Node(state)
                    search = R.AlphaZeroSearch(model=DeterministicEvaluator(kind), num_simulations=S, exploration_weight=c_puct)
                    pol, val = search.run(root)
                    bb0, bb1 = grid_to_bitboards(state.grid)
                    cN, cW, cP, ip = [0] * 7, [0.0] * 7, [0.0] * 7, [0.0] * 7
                    for a, ch in root.children.items():
                        cN[a.column], cW[a.column], cP[a.column] = ch.visit_count, ch.value_sum, ch.prior
                        ip[a.column] = pol[a]
                    cases.append(dict(
                        name=f"{name}_k{kind}_S{S}_c{c_puct}", bb0=bb0, bb1=bb1, player=state.player, S=S, c_puct=c_puct,
                        eval_kind=kind, legal=legal_mask(state), child_N=cN, child_W=cW, child_P=cP, root_W=root.value_sum,
                        root_N=root.visit_count, root_value=float(val), improved_policy=ip,
                        expected_move=sit.get("expected_move"), grid=full_grid(sit["grid"]),
                    ))
    return cases


def gen_selfplay(R):
    runs = []
    cfg = R.Config(6, 7, 4)
    specs = [
        dict(E=1, S=100, kind=UNIFORM, seed=0, init=None),
        dict(E=1, S=100, kind=HASH, seed=1, init=None),
        dict(E=8, S=50, kind=UNIFORM, seed=2, init=None),
        dict(E=8, S=100, kind=HASH, seed=3, init=None),
        dict(E=32, S=64, kind=HASH, seed=4, init=None),
        dict(E=64, S=200, kind=UNIFORM, seed=5, init=None),  # C2 in miniature (4096x200 is hours of CPython)
        dict(E=64, S=200, kind=HASH, seed=6, init=None),
        dict(E=4, S=100, kind=HASH, seed=7, init=FINAL_SITUATIONS[5]),  # non-empty initial state
        dict(E=3, S=800, kind=HASH, seed=8, init=None),
    ]
    for sp in specs:
        init_state = cfg.sample_initial_state() if sp["init"] is None else R.State.from_json(state_json(sp["init"]))
        gen = R.EpisodeGenerator(model=DeterministicEvaluator(sp["kind"]), num_simulations=sp["S"], num_episodes=sp["E"],
                                 game_initial_state=init_state)
        np.random.seed(sp["seed"])
        draws = [0]
        orig_choice = np.random.choice

        def counting_choice(*a, **k):  # counts draws only; the call goes to numpy unchanged
            draws[0] += 1
            return orig_choice(*a, **k)

        np.random.choice = counting_choice
        try:
            episodes = list(gen.generate_episodes())
        finally:
            np.random.choice = orig_choice
        after = float(np.random.random_sample())  # next value of the global stream after the run
        eps = []
        for ep in episodes:
            samples = []
            for s in ep.samples:
                bb0, bb1 = grid_to_bitboards(s.state.grid)
                pol = [0.0] * 7
                cnt = [0] * 7
                for a, p in s.policy.items():
                    pol[a.column] = p
                    cnt[a.column] = int(round(p * (sp["S"] - 1)))
                    assert cnt[a.column] / (sp["S"] - 1) == p
                samples.append(dict(bb0=bb0, bb1=bb1, player=s.state.player, counts=cnt, policy=pol))
            eps.append(dict(samples=samples, outcome=[float(v) for v in ep.samples[0].value]))
        ib0, ib1 = grid_to_bitboards(init_state.grid)
        runs.append(dict(E=sp["E"], S=sp["S"], eval_kind=sp["kind"], seed=sp["seed"], init_bb0=ib0, init_bb1=ib1,
                         init_player=init_state.player, n_draws=draws[0], next_uniform_after=after, episodes=eps))
        print(f"selfplay E={sp['E']} S={sp['S']} kind={sp['kind']}: {len(eps)} episodes, {draws[0]} draws", flush=True)
    return runs


def gen_rules(R):
    """Random playouts through the stand-in, each move cross-checked against connect4_game.py."""
    sys.path.insert(0, REFERENCE_SRC)
    from alphazero_simple.connect4_game import Connect4Game  # in-tree rules (row 0 = TOP, 0 empty, +-1)

    g4 = Connect4Game()
    rng = np.random.RandomState(1234)
    cfg = R.Config(6, 7, 4)
    games = []
    outcomes = {"p0": 0, "p1": 0, "draw": 0}
    for gi in range(300):
        st = cfg.sample_initial_state()
        board = g4.get_init_board()
        who = 1  # +1 == player 0
        plies = []
        bias = rng.rand(7) ** (3 if gi % 3 == 0 else 0.3)  # some column-biased games -> vertical wins & full columns
        while not st.has_ended:
            acts = st.actions
            cols = [a.column for a in acts]
            assert cols == sorted(cols)
            valid = g4.get_valid_moves(board)
            assert [c for c in range(7) if valid[c]] == cols, "legal moves disagree with connect4_game.py"
            w = bias[cols] + 1e-3
            a = acts[int(rng.choice(len(acts), p=w / w.sum()))]
            st = a.sample_next_state()
            board, who = g4.get_next_state(board, who, a.column)
            # encodings: in-tree grid is row 0 = top, +1 -> player 0, -1 -> player 1
            conv = np.where(board[::-1] == 0, -1, np.where(board[::-1] == 1, 0, 1))
            assert (conv == st.grid).all(), "board after move disagrees with connect4_game.py"
            r0 = g4.get_reward_for_player(board, 1)
            assert (r0 is not None) == st.has_ended
            if r0 is not None:
                assert [float(r0), float(-r0)] == st.reward.tolist(), "reward disagrees with connect4_game.py"
            bb0, bb1 = grid_to_bitboards(st.grid)
            plies.append(dict(col=a.column, bb0=bb0, bb1=bb1, player=st.player, legal=legal_mask(st),
                              ended=int(st.has_ended), reward=[int(v) for v in st.reward.tolist()]))
        rw = plies[-1]["reward"]
        outcomes["p0" if rw[0] > 0 else "p1" if rw[1] > 0 else "draw"] += 1
        games.append(plies)
    # forced draws: fill the board in a pattern with no 4-in-line
    # column order 0,1,2 filled with pattern that avoids wins is hard to hand-craft; search random fills instead
    tries = 0
    while outcomes["draw"] < 5 and tries < 200000:
        tries += 1
        st = cfg.sample_initial_state()
        plies = []
        while not st.has_ended:
            acts = st.actions
            a = acts[int(rng.randint(len(acts)))]
            st = a.sample_next_state()
            bb0, bb1 = grid_to_bitboards(st.grid)
            plies.append(dict(col=a.column, bb0=bb0, bb1=bb1, player=st.player, legal=legal_mask(st),
                              ended=int(st.has_ended), reward=[int(v) for v in st.reward.tolist()]))
        if plies[-1]["reward"] == [0, 0]:
            outcomes["draw"] += 1
            games.append(plies)
    print("rules playouts:", outcomes, "games:", len(games), flush=True)
    return dict(games=games, outcomes=outcomes)


def gen_nets(R):
    import torch

    out = {}
    states = [R.State.from_json(state_json(s)) for s in FINAL_SITUATIONS + START_SITUATIONS]
    enc = []
    for s in states:
        bb0, bb1 = grid_to_bitboards(s.grid)
        enc.append(dict(bb0=bb0, bb1=bb1, player=s.player))
    out["states"] = enc
    for name, cls in (("BasicNN", R.BasicNN), ("CNNModel", R.CNNModel)):
        torch.manual_seed(0)
        m = cls().eval()
        pol, val = m.predict(states)
        x = m._states_to_tensor(states)
        logits, v = m.forward(x)
        P = [[0.0] * 7 for _ in states]
        for i, d in enumerate(pol):
            for a, p in d.items():
                P[i][a.column] = p
        out[name] = dict(seed=0, priors=P, values=val, logits=logits.detach().tolist(), input_sum=float(x.sum()),
                         input_shape=list(x.shape), input_first=x[0].flatten().tolist())
    # C1: E=1, S=100, BasicNN, torch.manual_seed(0), np.random.seed(0)  (BASELINE.json configs[0])
    torch.manual_seed(0)
    torch.set_num_threads(1)
    model = R.BasicNN()
    gen = R.EpisodeGenerator(model=model, num_simulations=100, num_episodes=1,
                             game_initial_state=R.Config(6, 7, 4).sample_initial_state())
    np.random.seed(0)
    [ep] = list(gen.generate_episodes())
    samples = []
    for s in ep.samples:
        bb0, bb1 = grid_to_bitboards(s.state.grid)
        pol = [0.0] * 7
        for a, p in s.policy.items():
            pol[a.column] = p
        samples.append(dict(bb0=bb0, bb1=bb1, player=s.state.player, policy=pol))
    out["c1_basicnn_episode"] = dict(samples=samples, outcome=[float(v) for v in ep.samples[0].value])
    return out


def gen_training(R):
    """`Model.training_step` + `configure_optimizers` (models/base/model.py:27-48) and `format_dataset` (:76-82) of the reference on
    the fixture states: CE(soft visit targets) + MSE, Adam(lr 1e-3, weight decay 1e-4), four optimiser steps on one batch."""
    import torch

    states = [R.State.from_json(state_json(s)) for s in FINAL_SITUATIONS + START_SITUATIONS]
    policies, values = [], []
    for i, s in enumerate(states):
        acts = s.actions
        w = [1.0 + ((a.column * (i % 3 + 1)) % 5) for a in acts]
        policies.append({a: wi / sum(w) for a, wi in zip(acts, w)})
        values.append([1.0, -1.0] if i % 2 else [-1.0, 1.0])
    out = {}
    for name, cls in (("BasicNN", R.BasicNN), ("CNNModel", R.CNNModel)):
        torch.manual_seed(0)
        torch.set_num_threads(1)
        m = cls()
        ds = m.format_dataset(states, policies, values)
        x, p, v = ds.tensors
        m.train()
        opt = m.configure_optimizers()
        torch.manual_seed(1)  # Dropout masks of CNNModel
        losses = []
        for step in range(4):
            opt.zero_grad()
            loss = m.training_step((x, p, v), step)
            loss.backward()
            opt.step()
            losses.append(float(loss))
        out[name] = dict(losses=losses, policy_target=p.tolist(), value_target=v.tolist(), input_shape=list(x.shape),
                         param_abs_sum=float(sum(q.detach().double().abs().sum() for q in m.parameters())),
                         optimizer=dict(type=type(opt).__name__, lr=opt.defaults["lr"], weight_decay=opt.defaults["weight_decay"]))
    return out


def main():
    R = load_reference()
    os.makedirs(GOLDEN, exist_ok=True)
    meta = dict(generator="oracle/gen_golden.py", reference="pierreveron/alphazero-implementation @ /root/reference",
                game_layer="oracle/shims/simulator (stand-in; third-party simulator 0.0.4 source absent: PARITY UNPINNED)")
    for fname, fn in (("rules_goldens.json", gen_rules), ("search_goldens.json", gen_search),
                      ("nets_goldens.json", gen_nets), ("selfplay_goldens.json", gen_selfplay),
                      ("training_goldens.json", gen_training)):
        data = fn(R)
        with open(os.path.join(GOLDEN, fname), "w") as f:
            json.dump(dict(meta=meta, data=data), f, separators=(",", ":"))
        print("wrote", fname, os.path.getsize(os.path.join(GOLDEN, fname)), "bytes", flush=True)


if __name__ == "__main__":
    main()
