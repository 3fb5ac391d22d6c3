"""GPU: the hand-written tcgen05 evaluator for BasicNN (csrc/az_mlp.cu) vs plain PyTorch references of the same op."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import alphazero_implementation_b200 as az  # noqa: E402
from alphazero_implementation_b200.engine import LAYOUT_GRID_F32  # noqa: E402
from alphazero_implementation_b200.models import TensorCoreMLP  # noqa: E402


def _leaf_grids(n, seed=0):
    eng = az.Engine(num_games=n, num_simulations=16)
    eng.reset_games()
    eng.run_simulations(12, 2)
    u = torch.from_numpy(np.random.RandomState(seed).random_sample(n)).cuda()
    for _ in range(6):
        eng.sample_moves(u)
        eng.run_simulations(12, 2)
    eng.select_leaves()
    x = eng.gather_leaves(LAYOUT_GRID_F32).clone()
    eng.close()
    return x


def _bf16_emulation(m, x):
    """Same arithmetic as the kernel in PyTorch: bf16-rounded weights and activations, fp32 accumulation."""
    r = lambda t: t.to(torch.bfloat16).to(torch.float32)
    h = x.reshape(x.shape[0], -1)
    h = r(torch.relu(r(h) @ r(m.shared_layers[0].weight).T + m.shared_layers[0].bias))
    h = r(torch.relu(h @ r(m.shared_layers[2].weight).T + m.shared_layers[2].bias))
    logits = h @ r(m.policy_head.weight).T + m.policy_head.bias
    values = torch.tanh(h @ r(m.value_head[0].weight).T + m.value_head[0].bias)
    return logits, values


@pytest.mark.parametrize("n", [1, 127, 128, 129, 1000, 16384])
def test_fused_mlp_matches_references(n):
    torch.manual_seed(n)
    m = az.BasicNN().cuda().eval()
    with torch.no_grad():
        for p in m.parameters():  # random-init is tiny; scale up so that ReLU / tanh are exercised
            p.mul_(3.0)
    x = _leaf_grids(n, seed=n)
    mlp = TensorCoreMLP(m, torch.device("cuda", torch.cuda.current_device()))
    logits, values = mlp(x)
    torch.cuda.synchronize()
    with torch.no_grad():
        l_emu, v_emu = _bf16_emulation(m, x)
        l_ref, v_ref = m(x)  # plain fp32 module
    assert torch.isfinite(logits).all() and torch.isfinite(values).all()
    assert torch.allclose(logits, l_emu, atol=2e-3, rtol=1e-3), float((logits - l_emu).abs().max())
    assert torch.allclose(values, v_emu, atol=2e-3, rtol=1e-3), float((values - v_emu).abs().max())
    # against the fp32 module: bf16 operand rounding only
    assert float((logits - l_ref).abs().max()) < 5e-2 and float((values - v_ref).abs().max()) < 5e-2
    assert float((torch.softmax(logits, 1) - torch.softmax(l_ref, 1)).abs().max()) < 2e-2


def test_fused_mlp_in_the_search_loop():
    """BasicNN evaluated by the tcgen05 kernel inside AlphaZeroSearch: same trees as the cuBLAS bf16-free fp32 path up
    to the bf16 rounding of the evaluator (visit counts of the root children within a few visits)."""
    torch.manual_seed(3)
    model = az.BasicNN()
    roots = [az.Config().sample_initial_state()]
    out = []
    for dtype in (torch.float32, torch.bfloat16):
        s = az.AlphaZeroSearch(model=model, num_simulations=128, inference_dtype=dtype)
        assert (s._net.fused is not None) == (dtype == torch.bfloat16)
        nodes = [az.Node(r) for r in roots]
        s.run_simulations(nodes)
        out.append([ch.visit_count for ch in nodes[0].children.values()])
        assert sum(out[-1]) == 127
    assert max(abs(a - b) for a, b in zip(*out)) <= 8
    s.update_inference_model(model)  # weights are re-packed for the kernel
    nodes = [az.Node(r) for r in roots]
    s.run_simulations(nodes)
    assert [ch.visit_count for ch in nodes[0].children.values()] == out[1]


def test_fused_leaf_gather_equals_separate_gather():
    """az_mlp_forward_leaves (rows built from the leaf bitboards inside the kernel) == az_gather_leaves + az_mlp_forward,
    on the rows of the slots whose leaf waits for an evaluation (the fused kernel walks the engine's compacted list of those
    slots and leaves the other rows alone)."""
    torch.manual_seed(5)
    m = az.BasicNN().cuda().eval()
    n = 3000
    eng = az.Engine(num_games=n, num_simulations=64)
    eng.reset_games()
    u = torch.from_numpy(np.random.RandomState(1).random_sample(n)).cuda()
    for _ in range(14):  # deep enough that some leaves are terminal
        eng.run_simulations(40, 2)
        eng.sample_moves(u)
    eng.run_simulations(40, 2)
    eng.set_leaf_compaction(True)
    eng.select_leaves()
    status = eng.leaf_info()["status"]
    assert (status == 1).any() and (status == 0).any()
    mlp = TensorCoreMLP(m, torch.device("cuda", torch.cuda.current_device()))
    a_l, a_v = [t.clone() for t in mlp(eng.gather_leaves(LAYOUT_GRID_F32))]
    b_l, b_v = mlp.forward_leaves(eng)
    torch.cuda.synchronize()
    live = status == 0
    assert torch.equal(a_l[live], b_l[live]) and torch.equal(a_v[live], b_v[live])
    eng.close()
