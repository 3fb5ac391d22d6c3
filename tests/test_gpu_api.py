"""GPU: the reference-facing Python surface (AlphaZeroSearch / Node / EpisodeGenerator / Model.predict) read like
the reference's own usage and reproduce its recorded outputs."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import alphazero_implementation_b200 as az  # noqa: E402


def _state(case):
    return az.State(az.Config(6, 7, 4), case["bb0"], case["bb1"], case["player"])


def test_search_run_matches_reference(search_goldens):
    for case in [c for c in search_goldens if c["S"] == 100][:12]:
        ev = az.UniformEvaluator() if case["eval_kind"] == 1 else az.HashEvaluator()
        search = az.AlphaZeroSearch(model=ev, num_simulations=case["S"], exploration_weight=case["c_puct"])
        root = az.Node(_state(case))
        policy, value = search.run(root)  # ui/cli/player.py:60-64, notebooks/policy_comparison.ipynb#cell3
        assert [a.column for a in policy] == [c for c in range(7) if (case["legal"] >> c) & 1]
        assert [policy[a] for a in policy] == [case["improved_policy"][a.column] for a in policy]
        assert value == case["root_value"]
        assert root.visit_count == case["root_N"] and root.value_sum == case["root_W"]
        assert [ch.visit_count for ch in root.children.values()] == [case["child_N"][a.column] for a in root.children]
        assert [ch.prior for ch in root.children.values()] == [case["child_P"][a.column] for a in root.children]
        nxt = root.select_next_node()
        assert nxt.parent is root and not nxt.children and nxt.state.player == 1 - root.state.player


def test_win_in_one_fixture_is_found(search_goldens):
    case = next(c for c in search_goldens if c["name"].startswith("final0_k1_S100"))
    search = az.AlphaZeroSearch(model=az.UniformEvaluator(), num_simulations=100)
    root = az.Node(_state(case))
    policy, value = search.run(root)
    best = max(policy, key=policy.get)
    assert best.column == case["expected_move"] == 3
    assert root.children[best].value == 1.0  # winning child has Q = +1


def test_terminal_root_raises_like_the_reference():
    st = az.State(az.Config(), (1 << 0) | (1 << 7) | (1 << 14) | (1 << 21), (1 << 1) | (1 << 8) | (1 << 15), 1)
    assert st.has_ended and st.reward.tolist() == [1.0, -1.0] and st.actions == []
    with pytest.raises(AttributeError):
        az.AlphaZeroSearch(model=az.UniformEvaluator(), num_simulations=10).run(az.Node(st))


def test_full_tree_materialisation_is_consistent(search_goldens):
    case = next(c for c in search_goldens if c["name"].startswith("start3_k2_S300"))
    search = az.AlphaZeroSearch(model=az.HashEvaluator(), num_simulations=300)
    root = az.Node(_state(case))
    search.run_simulations([root], materialize="full")

    def walk(n):
        tot = 1
        if n.children:
            assert sum(ch.visit_count for ch in n.children.values()) == n.visit_count - 1
            for ch in n.children.values():
                assert ch.parent is n
                tot += walk(ch)
        return tot

    assert walk(root) > 300 and root.visit_count == 300


def test_state_action_interface():
    s = az.Config(6, 7, 4).sample_initial_state()
    assert s.player == 0 and not s.has_ended and [a.column for a in s.actions] == list(range(7))
    for col in (3, 3, 3, 3, 3, 3):
        s = s.actions[[a.column for a in s.actions].index(col)].sample_next_state()
    assert [a.column for a in s.actions] == [0, 1, 2, 4, 5, 6] and s.grid[:, 3].tolist() == [0, 1, 0, 1, 0, 1]
    with pytest.raises(ValueError):
        az.Action(s, 3).sample_next_state()
    assert az.State.from_json(s.to_json()) == s


@pytest.mark.parametrize("idx", [0, 1, 2, 3, 4, 7])
def test_episode_generator_matches_reference_transcript(selfplay_goldens, idx):
    run = selfplay_goldens[idx]
    ev = az.UniformEvaluator() if run["eval_kind"] == 1 else az.HashEvaluator()
    init = az.State(az.Config(), run["init_bb0"], run["init_bb1"], run["init_player"])
    gen = az.EpisodeGenerator(model=ev, num_simulations=run["S"], num_episodes=run["E"], game_initial_state=init)
    np.random.seed(run["seed"])
    episodes = list(gen.generate_episodes())
    assert len(episodes) == run["E"]
    for ep, ref in zip(episodes, run["episodes"]):
        assert len(ep) == len(ref["samples"])
        for s, r in zip(ep.samples, ref["samples"]):
            assert (s.state.bb0, s.state.bb1, s.state.player) == (r["bb0"], r["bb1"], r["player"])
            pol = [0.0] * 7
            for a, p in s.policy.items():
                pol[a.column] = p
            assert pol == r["policy"]
            assert s.value == ref["outcome"]
    # the global NumPy stream is left exactly where the reference leaves it
    assert float(np.random.random_sample()) == run["next_uniform_after"]
    d = episodes[0].to_dict()
    assert az.Episode.from_dict(d).samples[0].state == episodes[0].samples[0].state


def test_predict_matches_reference_nets(nets_goldens):
    """Model.predict (fp32) vs the reference's BasicNN / CNNModel outputs recorded with torch.manual_seed(0)."""
    states = [az.State(az.Config(), s["bb0"], s["bb1"], s["player"]) for s in nets_goldens["states"]]
    for name, cls in (("BasicNN", az.BasicNN), ("CNNModel", az.CNNModel)):
        torch.manual_seed(0)
        m = cls().eval()
        ref = nets_goldens[name]
        x = m._states_to_tensor(states)
        logits, _ = m.forward(x.cpu())
        assert np.allclose(logits.detach().numpy(), np.array(ref["logits"]), atol=1e-5), name  # same init, same arithmetic
        pol, val = m.predict(states)
        for i, d in enumerate(pol):
            assert [a.column for a in d] == [a.column for a in states[i].actions]
            got = [0.0] * 7
            for a, p in d.items():
                got[a.column] = p
            assert np.allclose(got, ref["priors"][i], atol=1e-3), (name, i)  # north_star tolerance 1e-3
            assert np.allclose(got, ref["priors"][i], atol=2e-6)
        assert np.allclose(np.array(val), np.array(ref["values"]), atol=2e-6)


def test_config1_basicnn_selfplay_policy_targets(nets_goldens):
    """BASELINE config 1: E=1, S=100, BasicNN (fp32), seeds 0 — policy targets within 1e-3 of the reference run,
    same positions and outcome."""
    torch.manual_seed(0)
    model = az.BasicNN()
    gen = az.EpisodeGenerator(model=model, num_simulations=100, num_episodes=1,
                              game_initial_state=az.Config(6, 7, 4).sample_initial_state())
    np.random.seed(0)
    [ep] = list(gen.generate_episodes())
    ref = nets_goldens["c1_basicnn_episode"]
    assert len(ep) == len(ref["samples"])
    for s, r in zip(ep.samples, ref["samples"]):
        assert (s.state.bb0, s.state.bb1, s.state.player) == (r["bb0"], r["bb1"], r["player"])
        pol = [0.0] * 7
        for a, p in s.policy.items():
            pol[a.column] = p
        assert np.allclose(pol, r["policy"], atol=1e-3)
    assert ep.samples[0].value == ref["outcome"]


def test_net_in_loop_graph_equals_eager_and_predict_path():
    """The CUDA-graphed step, the eager step and the generic `predict(states)` path build the same trees."""
    torch.manual_seed(1)
    model = az.BasicNN()
    roots = [az.Config().sample_initial_state()]
    s = roots[0]
    for c in (3, 2, 3):
        s = az.Action(s, c).sample_next_state()
        roots.append(s)
    out = []
    for kw in (dict(use_cuda_graph=True), dict(use_cuda_graph=False)):
        srch = az.AlphaZeroSearch(model=model, num_simulations=64, **kw)
        nodes = [az.Node(r) for r in roots]
        srch.run_simulations(nodes)
        out.append([[ch.visit_count for ch in n.children.values()] for n in nodes])

    class PredictOnly:  # duck-typed user evaluator: only `predict`
        def __init__(self, m):
            self.m = m.get_inference_clone().cuda()

        def get_inference_clone(self):
            return self

        def predict(self, states):
            return self.m.predict(states)

    srch = az.AlphaZeroSearch(model=PredictOnly(model), num_simulations=64)
    nodes = [az.Node(r) for r in roots]
    srch.run_simulations(nodes)
    out.append([[ch.visit_count for ch in n.children.values()] for n in nodes])
    assert out[0] == out[1]
    for a, b in zip(out[0], out[2]):  # softmax kernels differ in the last ulp at most; visit counts within 1
        assert max(abs(x - y) for x, y in zip(a, b)) <= 1


def test_resnet_bf16_inference_within_tolerance():
    """bf16 tensor-core inference vs the fp32 module on leaf positions: priors / values within 1e-2 here
    (random-init weights; the 1e-3 bar of north_star is for fp32 accumulate and is checked on the fp32 path)."""
    from alphazero_implementation_b200.models import InferenceNet

    torch.manual_seed(0)
    m = az.ResNet(num_res_blocks=2, num_channels=32).cuda().eval()
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.normal_(0, 0.1); mod.running_var.uniform_(0.5, 1.5)
    eng = az.Engine(num_games=256, num_simulations=8)
    eng.run_simulations(6, 2)
    eng.select_leaves()
    from alphazero_implementation_b200.engine import LAYOUT_PLANES_BF16, LAYOUT_PLANES_F32

    x32 = eng.gather_leaves(LAYOUT_PLANES_F32)
    with torch.no_grad():
        l_ref, v_ref = m(x32)
    f32 = InferenceNet(m, dtype=torch.float32)
    l32, v32 = f32(x32)
    assert torch.allclose(l32, l_ref, atol=1e-4) and torch.allclose(v32, v_ref, atol=1e-4)  # BN folding is exact to fp32 noise
    bf = InferenceNet(m, dtype=torch.bfloat16)
    l16, v16 = bf(eng.gather_leaves(LAYOUT_PLANES_BF16))
    assert torch.allclose(torch.softmax(l16, 1), torch.softmax(l_ref, 1), atol=1e-2)
    assert torch.allclose(v16, v_ref, atol=2e-2)
    eng.close()


def test_training_iteration_reduces_loss_and_syncs_weights():
    """Row (f1/f2): self-play -> replay buffer -> CE + MSE training step (Adam 1e-3, wd 1e-4) -> weight sync."""
    from alphazero_implementation_b200.trainer import Trainer

    torch.manual_seed(0)
    np.random.seed(0)
    model = az.BasicNN()
    tr = Trainer(model)
    hist = tr.train(num_iterations=3, episodes_per_iter=32, simulations_per_episode=32, epochs_per_iter=3,
                    initial_state=az.Config(6, 7, 4).sample_initial_state(), buffer_size=64)
    assert len(hist) == 3 and hist[0]["episodes"] == 32 and hist[1]["episodes"] == 64
    assert hist[0]["samples"] >= 32 * 7 and hist[2]["loss"] < hist[0]["loss"]
    assert hist[2]["graph_replays"] > 0 and hist[0]["optimizer_steps"] == 3 * -(-hist[0]["samples"] // 32)


def test_trainer_writes_the_reference_episode_files(tmp_path):
    """Row (f3): `save_every_n_iterations` -> `episodes_iter{N}.json` holding the replay deque in the reference's format
    (datamodule.py:71-80,109-112), readable by `load_episodes`, plus the weights of that iteration."""
    from alphazero_implementation_b200.episode import load_episodes
    from alphazero_implementation_b200.trainer import Trainer

    torch.manual_seed(0)
    np.random.seed(0)
    tr = Trainer(az.BasicNN())
    hist = tr.train(num_iterations=2, episodes_per_iter=16, simulations_per_episode=16, epochs_per_iter=1,
                    initial_state=az.Config(6, 7, 4).sample_initial_state(), buffer_size=24, save_every_n_iterations=2,
                    save_dir=str(tmp_path))
    assert sorted(p.name for p in tmp_path.iterdir()) == ["episodes_iter2.json", "model_iter2.ckpt"]
    eps = load_episodes(str(tmp_path / "episodes_iter2.json"))
    assert len(eps) == 24 == hist[1]["episodes"] and sum(len(e.samples) for e in eps) == hist[1]["samples"]
    s = eps[0].samples[0]
    assert abs(sum(s.policy.values()) - 1.0) < 1e-9 and s.value in ([1.0, -1.0], [-1.0, 1.0], [0.0, 0.0])
    ckpt = torch.load(tmp_path / "model_iter2.ckpt", weights_only=False)  # Lightning's file shape (trainer.py:66-70)
    assert set(ckpt["state_dict"].keys()) == set(az.BasicNN().state_dict().keys()) and "hyper_parameters" in ckpt
    back = az.BasicNN.load_from_checkpoint(str(tmp_path / "model_iter2.ckpt"))  # scripts/play.py:19
    assert all(torch.equal(a.cpu(), b.cpu()) for a, b in zip(back.state_dict().values(), tr.model.state_dict().values()))


def test_graphed_training_steps_equal_eager_steps():
    """Row (f1): the optimiser steps replayed from a CUDA graph are the eager steps - same minibatches, same weights."""
    from alphazero_implementation_b200.replay import ReplayBuffer
    from alphazero_implementation_b200.trainer import _GraphedTraining

    gen = az.EpisodeGenerator(model=az.UniformEvaluator(), num_simulations=24, num_episodes=64,
                              game_initial_state=az.Config().sample_initial_state())
    np.random.seed(3)
    rb = ReplayBuffer(buffer_size=64, num_simulations=24)
    for batch in gen.generate_batches(quota=64):
        rb.extend(batch)
    results = []
    for use_graph in (False, True):
        torch.manual_seed(11)
        model = az.BasicNN().cuda().train()
        opt = model.configure_optimizers()
        for group in opt.param_groups:
            group["capturable"] = True
        steps = _GraphedTraining(model, opt, batch_size=32, precision="32-true", device=torch.device("cuda", torch.cuda.current_device()))
        loss_sum, n = steps.run(rb, epochs=2, generator=torch.Generator().manual_seed(5), use_graph=use_graph)
        assert n == 2 * -(-rb.num_samples // 32)
        assert (steps.replays > 0) == use_graph
        results.append((float(loss_sum), [p.detach().clone() for p in model.parameters()]))
    (l0, w0), (l1, w1) = results
    assert l0 == pytest.approx(l1, rel=1e-5)
    for a, b in zip(w0, w1):
        assert torch.allclose(a, b, atol=1e-6, rtol=1e-5)


def test_replay_buffer_targets_match_reference_format(selfplay_goldens):
    """Dense policy / value targets built on the device equal the reference's format_dataset tensors."""
    from alphazero_implementation_b200.replay import ReplayBuffer

    run = selfplay_goldens[2]
    gen = az.EpisodeGenerator(model=az.UniformEvaluator(), num_simulations=run["S"], num_episodes=run["E"],
                              game_initial_state=az.Config().sample_initial_state())
    np.random.seed(run["seed"])
    rb = ReplayBuffer(buffer_size=5, num_simulations=run["S"])
    for batch in gen.generate_batches(quota=run["E"]):
        rb.extend(batch)
    assert len(rb) == 5  # deque(maxlen=buffer_size): only the last 5 episodes stay
    bb0, bb1, pl, policy, value = rb.tensors()
    ref_eps = run["episodes"][-5:]
    ref_pol = torch.tensor([s["policy"] for ep in ref_eps for s in ep["samples"]], dtype=torch.float32)
    ref_val = torch.tensor([ep["outcome"] for ep in ref_eps for _ in ep["samples"]], dtype=torch.float32)
    assert torch.allclose(policy.cpu(), ref_pol, atol=1e-7) and torch.equal(value.cpu(), ref_val)
    assert bb0.cpu().tolist() == [s["bb0"] for ep in ref_eps for s in ep["samples"]]
    x, p, v = next(rb.batches(az.BasicNN.input_layout, batch_size=8, shuffle=False))
    assert x.shape == (8, 6, 7) and p.shape == (8, 7) and v.shape == (8, 2)


def test_leaf_gather_into_pinned_host_batch_and_host_evaluator():
    """north_star (3): the leaf gather packs the board planes into a PINNED batch; a host-side evaluator reads it and its
    outputs go back through pinned buffers.  The search must equal the all-device path."""
    from alphazero_implementation_b200.engine import LAYOUT_GRID_F32, LAYOUT_PLANES_F32, POLICY_LOGITS

    E, S = 300, 24
    torch.manual_seed(5)
    net = az.BasicNN().eval()
    engs = [az.Engine(num_games=E, num_simulations=S) for _ in range(2)]
    for e in engs:
        e.reset_games()
    pinned_x = torch.empty((E, 6, 7), dtype=torch.float32).pin_memory()
    pinned_l = torch.empty((E, 7), dtype=torch.float32).pin_memory()
    pinned_v = torch.empty((E, 2), dtype=torch.float32).pin_memory()
    dev_net = az.BasicNN().cuda().eval()
    dev_net.load_state_dict(net.state_dict())
    for s in range(S):
        # device path
        engs[0].select_leaves()
        x = engs[0].gather_leaves(LAYOUT_GRID_F32)
        with torch.no_grad():
            l, v = dev_net(x)
        # pinned path: kernel writes the host batch, the host net reads it, the expand kernel reads host outputs
        engs[1].select_leaves()
        engs[1].gather_leaves(LAYOUT_GRID_F32, pinned_x)
        torch.cuda.synchronize()
        assert torch.equal(pinned_x, x.cpu())
        # feed both engines the SAME numbers so that the trees stay comparable bit for bit
        pinned_l.copy_(l.float().cpu())
        pinned_v.copy_(v.float().cpu())
        engs[0].expand_backup(l.float().contiguous(), v.float().contiguous(), POLICY_LOGITS)
        engs[1].expand_backup(pinned_l, pinned_v, POLICY_LOGITS)
        torch.cuda.synchronize()
    a, b = engs[0].root_stats(), engs[1].root_stats()
    for k in a:
        assert torch.equal(a[k], b[k]), k
    # the 3-plane layout lands in a pinned batch as well
    engs[1].set_roots(*[t.clone() for t in (engs[1].leaf_info()[k] for k in ("bb0", "bb1", "player"))])
    engs[1].select_leaves()
    px = torch.empty((E, 3, 6, 7), dtype=torch.float32).pin_memory()
    engs[1].gather_leaves(LAYOUT_PLANES_F32, px)
    torch.cuda.synchronize()
    assert torch.equal(px, engs[1].gather_leaves(LAYOUT_PLANES_F32).cpu())
    for e in engs:
        e.close()


def test_terminal_node_with_a_parent_is_visited_like_the_reference():
    """`run_simulations` on a node whose state has ended and that HAS a parent: the reference backs up the reward of the player
    who moved into it num_simulations times along the parent chain, without a sign flip at the terminal node (search.py:75-77,
    :48-57); a terminal node without a parent raises AttributeError."""
    st = az.State(az.Config(), (1 << 0) | (1 << 7) | (1 << 14), (1 << 1) | (1 << 8) | (1 << 15), 0)  # player 0 wins at column 3
    parent = az.Node(st)
    win = az.Node(az.Action(st, 3).sample_next_state(), parent=parent, prior=0.5)
    assert win.is_terminal and win.utility_values == [1.0, -1.0]
    fresh = az.Node(az.Config().sample_initial_state())
    search = az.AlphaZeroSearch(model=az.UniformEvaluator(), num_simulations=10)
    search.run_simulations([win, fresh])
    assert (win.visit_count, win.value_sum) == (10, 10.0)        # + reward[parent.player] per simulation
    assert (parent.visit_count, parent.value_sum) == (10, 10.0)  # no flip at the terminal node: the parent receives +v too
    assert fresh.visit_count == 10 and sum(c.visit_count for c in fresh.children.values()) == 9
