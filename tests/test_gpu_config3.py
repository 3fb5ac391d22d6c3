"""GPU: parity at the size of BASELINE configs[2] (16384 trees x 800 simulations, network in the loop).

1. The trees the production path builds (CUDA-graphed simulation steps, hand-written tcgen05 evaluator, in-kernel legal-only
   softmax) are BIT-EXACT with the oracle's when the oracle is fed the same evaluator outputs: a random sample of the
   16384 trees is searched again by oracle/c4_oracle.c (the reference algorithm, search.py:65-91) whose `predict` callback
   returns, for each leaf position, what the same kernels return for that position.
2. The evaluator outputs themselves (priors / values of `Connect4Model.predict`, models/games/connect4/model.py:19-43) are
   within north_star's 1e-3 of the fp32 reference `predict` (stock PyTorch on the CPU, oracle/net_eval.py) in the fp16 operand
   mode of the hand-written kernels; the bf16 mode's deviation is measured and bounded.
3. Policy targets (visit distributions at 800 simulations) and root values of >= 1k trees searched with the hand-written
   evaluator against the same trees searched with the fp32 evaluator.
"""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import alphazero_implementation_b200 as az  # noqa: E402
from alphazero_implementation_b200.engine import POLICY_LOGITS  # noqa: E402
from conftest import ROOT  # noqa: E402

S = 800


def random_positions(oracle, n, seed, max_plies=26):
    """n non-terminal positions after 0..max_plies random legal moves (host side, through the oracle's rules)."""
    rng = np.random.RandomState(seed)
    bb0, bb1, pl = np.zeros(n, np.uint64), np.zeros(n, np.uint64), np.zeros(n, np.uint8)
    target = rng.randint(0, max_plies + 1, size=n)
    legal = np.full(n, 0x7F, np.uint8)
    for ply in range(max_plies):
        go = np.nonzero(target > ply)[0]
        if not len(go):
            break
        # a random legal column per game
        r = rng.random_sample(len(go))
        cols = np.zeros(len(go), np.uint8)
        for j, i in enumerate(go):
            opts = [c for c in range(7) if (legal[i] >> c) & 1]
            cols[j] = opts[int(r[j] * len(opts))]
        nxt = oracle.env_step(bb0[go], bb1[go], pl[go], cols)
        ok = (nxt["status"] == 0) & (nxt["ended"] == 0)
        keep = go[ok]
        bb0[keep], bb1[keep], pl[keep], legal[keep] = nxt["bb0"][ok], nxt["bb1"][ok], nxt["player"][ok], nxt["legal"][ok]
        target[go[~ok]] = 0  # the move would end the game: this game stays where it is
    return bb0, bb1, pl, legal


class KernelEvaluator:
    """`predict` for the oracle: the production kernels applied to the given positions (a small helper engine whose roots are
    the positions: the first selection returns the root itself; expansion stores the legal-only softmax priors)."""

    def __init__(self, net, cap=64):
        self.net, self.cap = net, cap
        self.eng = az.Engine(num_games=cap, num_simulations=1)
        self.calls = 0

    def __call__(self, bb0, bb1, player, legal):
        n = len(bb0)
        assert n <= self.cap
        self.calls += 1
        e = self.eng
        e.set_roots(bb0, bb1, player)
        e.select_leaves()
        logits, values = self.net.forward_leaves(e)
        e.expand_backup(logits, values, POLICY_LOGITS)
        st = e.root_stats()
        return st["child_P"][:n].cpu().numpy(), values[:n].cpu().numpy()


@pytest.mark.parametrize("spec", ["resnet4x64", "basic"])
def test_config3_trees_bit_exact_with_oracle_fed_the_same_evaluator(oracle, spec):
    E, n_check = 16384, 24
    torch.manual_seed(0)
    model = az.ResNet(num_res_blocks=4, num_channels=64) if spec == "resnet4x64" else az.BasicNN()
    search = az.AlphaZeroSearch(model=model, num_simulations=S, inference_dtype=torch.bfloat16)
    bb0, bb1, pl, _ = random_positions(oracle, E, seed=11)
    eng = search.engine_for(E)
    eng.set_roots(bb0, bb1, pl)
    search.simulate(eng)  # the production path: graph-replayed (evaluator kernel, k_expand_select) x 800
    got = {k: v.cpu().numpy() for k, v in eng.root_stats().items()}
    assert (got["root_N"] == S).all() and (got["child_N"].sum(1) == S - 1).all()
    idx = np.random.RandomState(5).choice(E, n_check, replace=False)
    ev = KernelEvaluator(search._net)
    ref = oracle.search(bb0[idx], bb1[idx], pl[idx], S, py_eval=ev)
    assert ev.calls == S
    for k in ("child_N", "root_N"):
        assert (got[k][idx] == ref[k]).all(), k
    for k in ("child_W", "root_W", "child_P"):  # fp64 value sums and fp32 priors, bit for bit
        assert (got[k][idx] == ref[k]).all(), k
    search.close()


def _fp32_reference(model, bb0, bb1, pl, legal):
    from oracle.net_eval import TorchNetEvaluator

    return TorchNetEvaluator(model)(bb0, bb1, pl, legal)


def _kernel_outputs(model, dtype, bb0, bb1, pl, legal):
    from alphazero_implementation_b200.models import InferenceNet

    n = len(bb0)
    net = InferenceNet(model, dtype=dtype, device="cuda")
    assert net.evaluates_leaves_directly, "the hand-written kernel must be the one that runs"
    eng = az.Engine(num_games=n, num_simulations=1)
    eng.set_roots(bb0, bb1, pl)
    eng.select_leaves()
    logits, values = net.forward_leaves(eng)
    pri = eng.masked_softmax(logits, legal).cpu().numpy()
    val = values.cpu().numpy().copy()
    eng.close()
    return pri, val


def test_evaluator_outputs_within_1e_3_of_fp32_predict(oracle):
    """north_star: "policy targets and values must match within 1e-3 (fp32 accumulate)" against `Connect4Model.predict`."""
    n = 16384
    bb0, bb1, pl, legal = random_positions(oracle, n, seed=3)
    report = {}
    for name in ("resnet4x64", "basic"):
        torch.manual_seed(0)
        model = az.ResNet(num_res_blocks=4, num_channels=64) if name == "resnet4x64" else az.BasicNN()
        for mod in model.modules():  # non-trivial BatchNorm statistics, as after training
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.normal_(0, 0.1)
                mod.running_var.uniform_(0.5, 1.5)
                mod.weight.data.uniform_(0.5, 1.5)
                mod.bias.data.normal_(0, 0.1)
        p_ref, v_ref = _fp32_reference(model, bb0, bb1, pl, legal)
        for dname, dtype in (("fp16", torch.float16), ("bf16", torch.bfloat16)):
            p, v = _kernel_outputs(model, dtype, bb0, bb1, pl, legal)
            dp, dv = float(np.abs(p - p_ref).max()), float(np.abs(v - v_ref).max())
            report[f"{name}_{dname}"] = {"max_abs_dev_prior": dp, "max_abs_dev_value": dv, "positions": n}
            if dname == "fp16":
                assert dp <= 1e-3 and dv <= 1e-3, (name, dname, dp, dv)  # the tolerance north_star states
            else:
                assert dp <= 1e-3 and dv <= 5e-3, (name, dname, dp, dv)  # bf16: priors inside, values measured at ~1.5e-3
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    json.dump(report, open(os.path.join(out, "evaluator_deviation.json"), "w"), indent=1)


def test_policy_targets_at_800_sims_against_the_fp32_evaluator(oracle):
    """Visit distributions N_c / (N - 1) (`Node.improved_policy`, node.py:23-29) and root values of 1024 trees at 800 simulations:
    hand-written fp16 / bf16 evaluator in the loop against the fp32 module in the loop (cuDNN / cuBLAS with TF32 off).  A visit
    count is an integer: ONE simulation that goes to another child moves a policy target by 1/799 = 1.25e-3, so the bound that
    can hold for the max over 7168 targets is a few visits; the mean deviation is what is held against 1e-3."""
    E = 1024
    bb0, bb1, pl, _ = random_positions(oracle, E, seed=21)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    report = {}
    for name in ("resnet4x64", "basic"):
        torch.manual_seed(0)
        model = az.ResNet(num_res_blocks=4, num_channels=64) if name == "resnet4x64" else az.BasicNN()
        res = {}
        for dname, dtype in (("fp32", torch.float32), ("fp16", torch.float16), ("bf16", torch.bfloat16)):
            search = az.AlphaZeroSearch(model=model, num_simulations=S, inference_dtype=dtype)
            eng = search.engine_for(E)
            eng.set_roots(bb0, bb1, pl)
            search.simulate(eng)
            st = {k: v.cpu().numpy() for k, v in eng.root_stats().items()}
            res[dname] = (st["child_N"] / float(S - 1), st["root_W"] / st["root_N"])
            search.close()
        for dname in ("fp16", "bf16"):
            dp = np.abs(res[dname][0] - res["fp32"][0])
            dv = np.abs(res[dname][1] - res["fp32"][1])
            report[f"{name}_{dname}"] = {"trees": E, "policy_target_max_abs_dev": float(dp.max()), "policy_target_mean_abs_dev": float(dp.mean()),
                                         "trees_with_identical_visit_counts": int((dp.max(1) == 0).sum()),
                                         "root_value_max_abs_dev": float(dv.max()), "root_value_mean_abs_dev": float(dv.mean())}
        assert report[f"{name}_fp16"]["policy_target_mean_abs_dev"] <= 1e-3, report
        assert report[f"{name}_fp16"]["root_value_mean_abs_dev"] <= 1e-3, report
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    json.dump(report, open(os.path.join(out, "policy_target_deviation.json"), "w"), indent=1)
