"""GPU parity: the self-play loop (sample recording, np.random.choice-equivalent move sampling, slot recycling,
episode order) vs transcripts of the reference's EpisodeGenerator and vs the C oracle at config-2 size."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from conftest import golden_uniforms  # noqa: E402


def _run_engine(E, S, kind, uniforms_2d, init=(0, 0, 0), lanes=0, quota=None, fused=False):
    from alphazero_implementation_b200.engine import Engine

    quota = E if quota is None else quota
    eng = Engine(num_games=E, num_simulations=S, lanes_per_tree=lanes)
    eng.reset_games(*init)
    parts, count = [], 0
    for step in range(uniforms_2d.shape[0]):
        if fused:  # az_run_move_step: the same move step in one launch
            eng.run_move_step(S, kind, torch.from_numpy(uniforms_2d[step]).cuda())
        else:
            eng.run_simulations(S, kind)
            eng.sample_moves(torch.from_numpy(uniforms_2d[step]).cuda())
        ne, _ = eng.episode_counts()
        if ne:
            parts.append(eng.drain_episodes())
            count += ne
            if count >= quota:
                break
    stats = eng.stats()
    eng.close()
    return parts, stats, step + 1


def _flatten(parts, quota):
    eps = []
    for b in parts:
        for e in range(len(b)):
            o, n = int(b.ep_offset[e]), int(b.ep_len[e])
            eps.append(dict(slot=int(b.ep_slot[e]), step=int(b.ep_step[e]), outcome=b.ep_outcome[e].tolist(),
                            bb0=b.s_bb0[o:o + n].tolist(), bb1=b.s_bb1[o:o + n].tolist(), player=b.s_player[o:o + n].tolist(),
                            counts=b.s_counts[o:o + n].tolist()))
    return eps[:quota]


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("lanes", [8, 32])
def test_selfplay_goldens(selfplay_goldens, lanes, fused):
    for run in selfplay_goldens:
        E, S = run["E"], run["S"]
        u = golden_uniforms(run)
        steps = (run["n_draws"] + E - 1) // E + 1
        u2 = u[: steps * E].reshape(steps, E)
        parts, stats, _ = _run_engine(E, S, run["eval_kind"], u2, init=(run["init_bb0"], run["init_bb1"], run["init_player"]), lanes=lanes,
                                     fused=fused)
        got = _flatten(parts, E)
        assert len(got) == len(run["episodes"]) == E
        for g, ep in zip(got, run["episodes"]):
            assert g["outcome"] == [int(v) for v in ep["outcome"]]
            assert g["bb0"] == [s["bb0"] for s in ep["samples"]] and g["bb1"] == [s["bb1"] for s in ep["samples"]]
            assert g["player"] == [s["player"] for s in ep["samples"]]
            assert g["counts"] == [s["counts"] for s in ep["samples"]]
            assert [[c / (S - 1) for c in row] for row in g["counts"]] == [s["policy"] for s in ep["samples"]]
        # draws consumed up to the quota slot of the last step == the reference's np.random.choice calls
        last = got[-1]
        assert last["step"] * E + last["slot"] + 1 == run["n_draws"]


@pytest.mark.parametrize("E,S,kind,lanes,fused", [(4096, 200, 1, 8, False), (4096, 200, 1, 8, True), (4096, 200, 2, 32, False), (1024, 64, 2, 8, True),
                                                  (4096, 200, 2, 16, True), (333, 37, 1, 32, True)])
def test_selfplay_vs_oracle_config2(oracle, E, S, kind, lanes, fused):
    """BASELINE config 2 (4096 games x 200 sims): a whole self-play round until E episodes, all per-move visit
    counts, positions, outcomes and the episode order bit-exact against the C oracle."""
    rng = np.random.RandomState(E + S)
    u2 = rng.random_sample((60, E))
    ref = oracle.selfplay(E, S, u2, eval_kind=kind)
    parts, stats, steps = _run_engine(E, S, kind, u2, lanes=lanes, fused=fused)
    got = _flatten(parts, E)
    assert len(got) == E == len(ref.ep_slot) and steps == ref.n_steps
    assert [g["slot"] for g in got] == ref.ep_slot.tolist() and [g["step"] for g in got] == ref.ep_step.tolist()
    assert [g["outcome"] for g in got] == ref.ep_outcome.tolist()
    assert np.array([len(g["bb0"]) for g in got]).tolist() == ref.ep_len.tolist()
    assert (np.concatenate([np.array(g["bb0"], np.uint64) for g in got]) == ref.s_bb0).all()
    assert (np.concatenate([np.array(g["bb1"], np.uint64) for g in got]) == ref.s_bb1).all()
    assert (np.concatenate([np.array(g["counts"], np.int32) for g in got]) == ref.s_counts).all()
    assert (np.concatenate([np.array(g["player"], np.uint8) for g in got]) == ref.s_player).all()
    assert stats["simulations"] == ref.n_sims and stats["evaluations"] == ref.n_evals
    outcomes = np.array([g["outcome"][0] for g in got])
    assert (outcomes == 1).sum() > 0 and (outcomes == -1).sum() > 0


@pytest.mark.parametrize("kind", [1, 2])
def test_fused_move_step_equals_two_calls(kind):
    """az_run_move_step == az_run_simulations + az_sample_moves: finished flags, next roots, move counters, ring."""
    import alphazero_implementation_b200 as az

    E, S = 777, 48
    u = torch.from_numpy(np.random.RandomState(3).random_sample((30, E))).cuda()
    engs = [az.Engine(num_games=E, num_simulations=S + 1) for _ in range(2)]  # + 1: the root probe below is a simulation
    fin = [torch.full((E,), 7, dtype=torch.uint8, device="cuda") for _ in range(2)]
    for e in engs:
        e.reset_games()
    for step in range(30):
        engs[0].run_simulations(S, kind)
        engs[0].sample_moves(u[step], fin[0])
        engs[1].run_move_step(S, kind, u[step], fin[1])
        assert torch.equal(fin[0], fin[1])
        # the next move's search starts from the same roots: the first selection of a fresh tree returns the root itself
        for e in engs:
            e.select_leaves()
        a, b = engs[0].leaf_info(), engs[1].leaf_info()
        for k in a:
            assert torch.equal(a[k], b[k]), (step, k)
    sa, sb = engs[0].stats(), engs[1].stats()
    assert sa["moves"] == sb["moves"] and sa["episodes"] == sb["episodes"] and sa["episodes"] > 0
    da, db = engs[0].drain_episodes(), engs[1].drain_episodes()
    for f in ("ep_slot", "ep_step", "ep_len", "ep_outcome", "s_bb0", "s_bb1", "s_player", "s_counts"):
        assert (getattr(da, f) == getattr(db, f)).all(), f
    for e in engs:
        e.close()


def test_sharded_selfplay_equals_unsharded():
    """games never interact, so a rank that owns slots [lo, hi) and is fed those slots' uniforms must play
    exactly the games the single-GPU run plays in those slots (here: two half-size engines against one full-size engine)."""
    import alphazero_implementation_b200 as az
    from alphazero_implementation_b200.distributed import shard_range

    E, S, steps = 1000, 64, 45
    u = np.random.RandomState(11).random_sample((steps, E))

    def play(lo, hi):
        eng = az.Engine(num_games=hi - lo, num_simulations=S)
        eng.reset_games()
        out = {}
        for s in range(steps):
            eng.run_move_step(S, 2, torch.from_numpy(np.ascontiguousarray(u[s, lo:hi])).cuda())
            if s % 8 == 7 or s == steps - 1:  # the ring holds 2 * num_games + 64 episodes
                b = eng.drain_episodes()
                for e in range(len(b)):
                    o, n = int(b.ep_offset[e]), int(b.ep_len[e])
                    out[(int(b.ep_step[e]), int(b.ep_slot[e]) + lo)] = (b.ep_outcome[e].tolist(), b.s_bb0[o:o + n].tolist(),
                                                                       b.s_bb1[o:o + n].tolist(), b.s_counts[o:o + n].tolist())
        eng.close()
        return out

    full = play(0, E)
    merged = {}
    for r in range(2):
        merged.update(play(*shard_range(E, r, 2)))
    assert len(full) > E // 2 and full.keys() == merged.keys()
    for k in full:
        assert full[k] == merged[k], k
