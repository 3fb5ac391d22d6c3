"""Host-side logic that needs no GPU: episode ordering, JSON shapes, grid <-> bitboard formatting."""
import json

import numpy as np

from alphazero_implementation_b200.engine import sort_episode_batch
from alphazero_implementation_b200.episode import Episode, Sample
from alphazero_implementation_b200.game import Action, Config, State, bitboards_to_grid, grid_to_bitboards


def test_grid_bitboard_round_trip(rules_goldens):
    for plies in rules_goldens["games"][:50]:
        for p in plies:
            g = bitboards_to_grid(p["bb0"], p["bb1"])
            assert grid_to_bitboards(g) == (p["bb0"], p["bb1"])
            assert (g >= 0).sum() == bin(p["bb0"] | p["bb1"]).count("1")


def test_state_json_shape_matches_reference_fixture():
    # notebooks/episode_generation_testing.ipynb#cell2
    d = {"config": {"count": 4, "height": 6, "width": 7},
         "grid": [[0, 0, 0, -1, -1, 1, -1], [-1] * 7, [-1] * 7, [-1] * 7, [-1] * 7, [-1] * 7], "player": 0}
    s = State.from_json(d)
    assert s.to_json() == d
    assert s.bb0 == (1 << 0) | (1 << 7) | (1 << 14) and s.bb1 == 1 << 35


def test_episode_json_round_trip():
    cfg = Config(6, 7, 4)
    s = State(cfg, 1, 1 << 7, 0, legal=0x7F, ended=False, reward=(0, 0))
    pol = {Action(s, c): (c + 1) / 28 for c in range(7)}
    ep = Episode()
    ep.add_sample(Sample(s, pol, [0.0, 0.0]))
    ep.backpropagate_outcome([1.0, -1.0])
    d = json.loads(json.dumps(ep.to_dict()))
    assert list(d["samples"][0]["policy"].keys())[0] == "{'column': 0}"  # str(action.to_json()), episode.py:22
    back = Episode.from_dict(d)
    assert back.samples[0].state == s and back.samples[0].value == [1.0, -1.0]
    assert [a.column for a in back.samples[0].policy] == list(range(7))
    assert list(back.samples[0].policy.values()) == list(pol.values())


def test_sort_episode_batch_orders_by_step_then_slot():
    d = dict(ep_slot=np.array([5, 1, 3], np.int32), ep_step=np.array([2, 2, 1], np.int32), ep_len=np.array([2, 1, 3], np.int32),
             ep_offset=np.array([0, 2, 3], np.int64), ep_outcome=np.array([[1, -1], [0, 0], [-1, 1]], np.int8),
             s_bb0=np.arange(6, dtype=np.int64), s_bb1=np.arange(6, dtype=np.int64) * 10, s_player=np.arange(6, dtype=np.uint8),
             s_counts=np.arange(42, dtype=np.int32).reshape(6, 7))
    b = sort_episode_batch(d)
    assert b.ep_slot.tolist() == [3, 1, 5] and b.ep_step.tolist() == [1, 2, 2]
    assert b.ep_offset.tolist() == [0, 3, 4] and b.ep_len.tolist() == [3, 1, 2]
    assert b.s_bb0.tolist() == [3, 4, 5, 2, 0, 1]
    assert b.ep_outcome.tolist() == [[-1, 1], [0, 0], [1, -1]]
    assert b.num_samples == 6 and len(b) == 3


def test_episode_file_round_trip(tmp_path):
    from alphazero_implementation_b200.episode import load_episodes, save_episodes

    cfg = Config(6, 7, 4)
    eps = []
    for k in range(3):
        s = State(cfg, 1 << (7 * k), 1 << (7 * k + 1), 0, legal=0x7F, ended=False, reward=(0, 0))
        ep = Episode()
        ep.add_sample(Sample(s, {Action(s, c): 1 / 7 for c in range(7)}, [0.0, 0.0]))
        ep.backpropagate_outcome([-1.0, 1.0])
        eps.append(ep)
    path = tmp_path / "episodes_iter1.json"
    save_episodes(eps, str(path))
    raw = json.load(open(path))
    assert isinstance(raw, list) and set(raw[0]) == {"samples"} and set(raw[0]["samples"][0]) == {"state", "policy", "value"}
    back = load_episodes(str(path))
    assert [e.samples[0].state for e in back] == [e.samples[0].state for e in eps]
    assert back[2].samples[0].value == [-1.0, 1.0]


def _fake_batch(rng, n_eps, first_id=0, shuffled=False):
    """Episodes whose samples carry recognisable values: bb0 = 1000 * episode id + ply."""
    import torch

    lens = rng.randint(1, 9, size=n_eps)
    order = rng.permutation(n_eps) if shuffled else np.arange(n_eps)  # storage order of the sample blocks
    offs = np.zeros(n_eps, np.int64)
    pos = 0
    for e in order:
        offs[e] = pos
        pos += lens[e]
    bb0 = np.zeros(pos, np.int64)
    counts = np.zeros((pos, 7), np.int32)
    for e in range(n_eps):
        for q in range(lens[e]):
            bb0[offs[e] + q] = 1000 * (first_id + e) + q
            counts[offs[e] + q] = (first_id + e + q) % 5
    outcome = np.stack([np.where(np.arange(n_eps) % 3 == 0, 1, -1), -np.where(np.arange(n_eps) % 3 == 0, 1, -1)], 1).astype(np.int8)
    d = dict(ep_len=lens.astype(np.int32), ep_offset=offs, ep_outcome=outcome, s_bb0=bb0, s_bb1=bb0 + 7, s_player=(bb0 % 2).astype(np.uint8),
             s_counts=counts)
    return {k: torch.from_numpy(v) for k, v in d.items()}, lens


def test_replay_buffer_is_a_deque_of_episodes():
    """datamodule.py:57 `deque(maxlen=buffer_size)` of episodes, flattened per datamodule.py:114-122 - on flat chunked storage."""
    from collections import deque

    import torch

    from alphazero_implementation_b200.replay import ReplayBuffer

    rng = np.random.RandomState(0)
    rb = ReplayBuffer(buffer_size=25, num_simulations=11, device="cpu")
    ref = deque(maxlen=25)  # (episode id, length, outcome row)
    next_id = 0
    for n_eps, shuffled in ((10, False), (7, True), (12, False), (30, True), (1, False), (0, False)):
        batch, lens = _fake_batch(rng, n_eps, first_id=next_id, shuffled=shuffled)
        rb.extend(batch)
        for e in range(n_eps):
            ref.append((next_id + e, int(lens[e]), batch["ep_outcome"][e].tolist()))
        next_id += n_eps
        assert len(rb) == len(ref) and rb.num_samples == sum(l for _, l, _ in ref)
        bb0, bb1, pl, policy, value = rb.tensors()
        want_bb0 = [1000 * i + q for i, l, _ in ref for q in range(l)]
        assert bb0.tolist() == want_bb0 and bb1.tolist() == [v + 7 for v in want_bb0]
        assert value.tolist() == [[float(o[0]), float(o[1])] for _, l, o in ref for _ in range(l)]
        want_counts = torch.tensor([[(i + q) % 5] * 7 for i, l, _ in ref for q in range(l)], dtype=torch.float32)
        assert torch.equal(policy, want_counts / 10.0)  # improved_policy = N_c / (S - 1)
        # the same content as flat host arrays (what the episodes_iter{N}.json writer iterates over)
        eb = rb.to_episode_batch()
        assert len(eb) == len(ref) and eb.ep_len.tolist() == [l for _, l, _ in ref]
        assert eb.ep_offset.tolist() == [sum(l for _, l, _ in list(ref)[:k]) for k in range(len(ref))]
        assert eb.s_bb0.tolist() == want_bb0 and eb.ep_outcome.tolist() == [o for _, _, o in ref]
        assert eb.s_counts.dtype == np.int32 and eb.s_counts.shape == (len(want_bb0), 7)


def test_checkpoint_round_trip_in_the_reference_file_shape(tmp_path):
    """`Model.load_from_checkpoint` (scripts/play.py:19-25) reads a Lightning-shaped checkpoint (`state_dict` +
    `hyper_parameters`, what the reference's ModelCheckpoint writes, trainer.py:66-70) and a bare state dict; stale keyword
    arguments of the reference's call site (`height=`, `width=`, ...) that the constructor does not take are ignored."""
    import torch

    from alphazero_implementation_b200.models import BasicNN, CNNModel, ResNet

    torch.manual_seed(3)
    for cls, kw in ((BasicNN, {}), (CNNModel, {}), (ResNet, dict(num_res_blocks=2, num_channels=32))):
        m = cls(**kw)
        p = tmp_path / f"{cls.__name__}.ckpt"
        m.save_checkpoint(str(p), epoch=7, global_step=123)
        ck = torch.load(p, weights_only=False)
        assert set(ck) >= {"state_dict", "hyper_parameters", "epoch", "global_step"} and ck["hyper_parameters"]["model_name"] == cls.__name__
        back = cls.load_from_checkpoint(str(p), height=6, width=7, max_actions=7, num_players=2)
        assert not back.training
        assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), back.state_dict().values()))
    # a checkpoint written by the reference's Lightning run: same keys, plus trainer state we do not need
    ref_shaped = {"state_dict": BasicNN().state_dict(), "hyper_parameters": {"learning_rate": 1e-3, "model_name": "BasicNN"},
                  "epoch": 249, "global_step": 62820, "optimizer_states": [], "lr_schedulers": [], "loops": {}}
    torch.save(ref_shaped, tmp_path / "ref.ckpt")
    assert isinstance(BasicNN.load_from_checkpoint(str(tmp_path / "ref.ckpt")), BasicNN)
    torch.save(BasicNN().state_dict(), tmp_path / "bare.pt")
    assert isinstance(BasicNN.load_from_checkpoint(str(tmp_path / "bare.pt")), BasicNN)


def test_shard_table_is_validated_before_any_collective():
    from alphazero_implementation_b200.distributed import shard_range

    assert [shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    shards = [shard_range(8, r, 4, trainer_share=1.0) for r in range(4)]
    assert shards[0] == (0, 8) and all(lo == hi for lo, hi in shards[1:])  # Trainer.train raises on EVERY rank for such a table


def test_training_step_losses_match_the_reference(nets_goldens):
    """Row f1: `Model.training_step` + `configure_optimizers` (models/base/model.py:27-48) - CE(soft visit targets) + MSE,
    Adam(lr 1e-3, weight decay 1e-4) - reproduce the losses of four optimiser steps the REFERENCE's own classes took on the same
    batch (tests/golden/training_goldens.json, oracle/gen_golden.py:gen_training).  CPU, fp32, same seeds; CNNModel includes
    train-mode BatchNorm and Dropout(0.3)."""
    import torch

    from alphazero_implementation_b200.models import BasicNN, CNNModel
    from conftest import load_golden
    from oracle.net_eval import grid_f32, planes_f32

    tg = load_golden("training_goldens.json")
    st = nets_goldens["states"]
    bb0 = np.array([s["bb0"] for s in st], np.uint64)
    bb1 = np.array([s["bb1"] for s in st], np.uint64)
    pl = np.array([s["player"] for s in st], np.uint8)
    for name, cls in (("BasicNN", BasicNN), ("CNNModel", CNNModel)):
        ref = tg[name]
        torch.manual_seed(0)
        torch.set_num_threads(1)
        m = cls()
        x = torch.from_numpy(grid_f32(bb0, bb1) if name == "BasicNN" else planes_f32(bb0, bb1, pl))
        assert list(x.shape) == ref["input_shape"]
        p, v = torch.tensor(ref["policy_target"]), torch.tensor(ref["value_target"])
        m.train()
        opt = m.configure_optimizers()
        assert type(opt).__name__ == ref["optimizer"]["type"] and opt.defaults["lr"] == ref["optimizer"]["lr"]
        assert opt.defaults["weight_decay"] == ref["optimizer"]["weight_decay"]
        torch.manual_seed(1)
        losses = []
        for step in range(4):
            opt.zero_grad()
            loss = m.training_step((x, p, v), step)
            loss.backward()
            opt.step()
            losses.append(float(loss.detach()))
        assert np.allclose(losses, ref["losses"], rtol=2e-5, atol=1e-6), (name, losses, ref["losses"])
        s = float(sum(q.detach().double().abs().sum() for q in m.parameters()))
        assert abs(s - ref["param_abs_sum"]) <= 1e-6 * ref["param_abs_sum"], name


def test_elo_update_is_the_notebooks():
    """src/elo.ipynb#cell1: K = 32, ratings truncated to int."""
    from alphazero_implementation_b200.player import calculate_expected_score, update_elo

    assert calculate_expected_score(1500, 1500) == 0.5
    assert update_elo(1500, 1500, 1.0) == (1516, 1484)
    assert update_elo(1500, 1500, 0.5) == (1500, 1500)
    assert update_elo(1516, 1484, 0.0) == (int(1516 + 32 * (0 - 1 / (1 + 10 ** (-32 / 400)))), int(1484 + 32 * (1 - (1 - 1 / (1 + 10 ** (-32 / 400))))))


def test_player_move_rule_on_visit_counts():
    """ui/cli/player.py:66-74 on root child visit counts: temperature 0 = first maximum of the visit policy, t = p ** (1 / t)
    renormalised and sampled, inf = a uniformly random legal column (host logic of `AlphaZeroPlayer` / `Arena`)."""
    import random

    from alphazero_implementation_b200.player import _pick

    counts = np.array([10, 40, 0, 40, 5, 0, 4])
    legal = 0b1011011  # columns 0, 1, 3, 4, 6
    rng = random.Random(0)
    assert _pick(counts, legal, 0, rng) == 1  # first of the two maxima, like max() over the policy dict in action order
    assert {_pick(counts, legal, float("inf"), rng) for _ in range(200)} == {0, 1, 3, 4, 6}
    picks = [_pick(counts, legal, 1.0, rng) for _ in range(4000)]
    freq = np.bincount(picks, minlength=7) / 4000
    assert np.allclose(freq, counts / counts.sum(), atol=0.03) and freq[2] == 0 and freq[5] == 0
    cold = [_pick(counts, legal, 0.05, rng) for _ in range(200)]
    assert set(cold) <= {1, 3}  # (40 / 99) ** 20 against (10 / 99) ** 20: the maxima take everything


def test_packed_weights_of_the_fused_filter_row_kernel():
    """csrc/az_resnet_wide.cu reads (a) the layer-pipelined kernel's trunk pieces [9 taps][64 out][16 in] as [3 filter rows][192 =
    (filter column, out)][16 in] - the canonical K-major order is row-group-major, so the two views are the same bytes - and (b) head
    pieces [3 filter rows][64][16]: rows 0..31 = policy conv1x1 (centre tap only), row 32 + 8 kx + v = value channel v, filter column kx.
    Unpack both and compare with the BatchNorm-folded convolution weights."""
    import torch

    import alphazero_implementation_b200 as az
    from alphazero_implementation_b200.models import _fold_bn, pack_head_weights, pack_trunk_weights_pipe

    torch.manual_seed(0)
    m = az.ResNet(num_res_blocks=1, num_channels=64).eval()
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.normal_(0, 0.2)
            mod.running_var.uniform_(0.5, 1.5)
    w, b = pack_trunk_weights_pipe(m, "cpu", torch.float32)
    piece = 9 * 64 * 16
    assert w.numel() == piece * (1 + 2 * 4) and b.shape == (3, 64)
    w1, _ = _fold_bn(m.residual_blocks[0].conv1, m.residual_blocks[0].bn1)
    for ks in range(4):  # piece 1 + ks = K chunk ks of conv1, viewed as [3 dy][192][16]: [N/8][K/8][8][8] per filter row
        p = w[piece * (1 + ks):piece * (2 + ks)].reshape(3, 24, 2, 8, 8).permute(0, 1, 3, 2, 4).reshape(3, 192, 16)
        for ky in range(3):
            for kx in range(3):
                assert torch.equal(p[ky, 64 * kx:64 * kx + 64], w1[:, 16 * ks:16 * ks + 16, ky, kx])
    conv, hb, *_ = pack_head_weights(m, "cpu", torch.float32, wide=True)
    assert conv.numel() == 4 * 3 * 64 * 16 and hb.shape == (48,)
    rows = conv.reshape(4, 3, 8, 2, 8, 8).permute(1, 2, 4, 0, 3, 5).reshape(3, 64, 64)  # [ky][row][in]
    wp, bp = _fold_bn(m.policy_head[0], m.policy_head[1])
    wv, bv = _fold_bn(m.value_head[0], m.value_head[1])
    assert torch.equal(rows[1, :32], wp[:, :, 0, 0]) and not rows[0, :32].any() and not rows[2, :32].any()
    for ky in range(3):
        for kx in range(3):
            assert torch.equal(rows[ky, 32 + 8 * kx:35 + 8 * kx], wv[:, :, ky, kx])
            assert not rows[ky, 35 + 8 * kx:40 + 8 * kx].any()
    assert not rows[:, 56:].any() and torch.equal(hb[:32], bp) and torch.equal(hb[32:35], bv)
