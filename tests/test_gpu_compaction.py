"""GPU: the list of the leaves that wait for an evaluation - written by one more launch per selection (k_compact_leaves, ascending)
or by k_expand_select itself (az_set_leaf_compaction(h, 2): one launch less per simulation, ticket order) - and the search on top."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import alphazero_implementation_b200 as az  # noqa: E402
from alphazero_implementation_b200.engine import POLICY_LOGITS  # noqa: E402


def _mid_game_engine(n, sims, seed):
    eng = az.Engine(num_games=n, num_simulations=sims)
    eng.reset_games()
    u = torch.from_numpy(np.random.RandomState(seed).random_sample(n)).cuda()
    for _ in range(13):  # deep enough that many leaves are terminal and some games restart
        eng.run_simulations(24, 2)
        eng.sample_moves(u)
    return eng


def _hash_outputs(eng):
    """A deterministic evaluator on the leaves (a function of the leaf position only), as dense [n, 7] / [n, 2] rows."""
    info = eng.leaf_info()
    h = (info["bb0"] * 0x9E3779B97F4A7C15 + info["bb1"] * 0x7F4A7C15 + info["player"].to(torch.int64)) & 0xFFFFFFFF
    logits = torch.stack([((h >> (3 * c)) & 31).to(torch.float32) / 8.0 for c in range(7)], 1).contiguous()
    v = (((h >> 21) & 255).to(torch.float32) - 128.0) / 128.0
    return logits, torch.stack([v, -v], 1).contiguous()


@pytest.mark.parametrize("n", [5, 777, 16384, 20000])
def test_fused_list_is_a_permutation_of_the_ordered_list(n):
    engs = []
    for fused in (False, True):
        eng = _mid_game_engine(n, 64, seed=n)
        eng.set_leaf_compaction(True, fused=fused)
        eng.select_leaves()
        lists = []
        for _ in range(6):
            logits, values = _hash_outputs(eng)
            eng.expand_backup_select(logits, values, POLICY_LOGITS)
            lst, cnt = eng.leaf_compact()
            torch.cuda.synchronize()
            c = cnt.cpu()
            assert int(c[2]) == 0 and int(c[3]) == 0  # the ticket word is back to zero between launches
            k = int(c[0])
            status = eng.leaf_info()["status"]
            assert k == int((status == 0).sum())
            lists.append(lst[:k].clone())
        engs.append((eng, lists))
    (e0, l0), (e1, l1) = engs
    for a, b in zip(l0, l1):
        assert torch.equal(a, a.sort().values)  # the extra launch writes the list in ascending order
        assert torch.equal(a, b.sort().values)  # the fused one writes the same slots
    s0, s1 = e0.root_stats(), e1.root_stats()
    for key in ("child_N", "child_W", "child_P", "root_W", "root_N"):
        assert torch.equal(s0[key], s1[key]), key
    e0.close(); e1.close()


def test_search_with_a_network_is_the_same_with_either_list(monkeypatch):
    """The production loop (CUDA graph of k_resnet_wide + k_expand_select) with the fused list against the loop with the extra
    launch: every tree's root statistics bit-identical (a position's outputs do not depend on the batch it falls into)."""
    from alphazero_implementation_b200.search import AlphaZeroSearch

    torch.manual_seed(3)
    model = az.ResNet(num_res_blocks=2, num_channels=64).cuda().eval()
    n, S = 3000, 96
    outs = []
    for fused in ("0", "1"):
        monkeypatch.setenv("AZ_COMPACT_FUSED", fused)
        search = AlphaZeroSearch(model=model, num_simulations=S, inference_dtype=torch.bfloat16)
        eng = search.engine_for(n, exact=True)
        eng.reset_games()
        u = torch.from_numpy(np.random.RandomState(2).random_sample((9, n))).cuda()
        for i in range(9):
            search.simulate_and_move(eng, u[i])
        search.simulate(eng)  # the tenth search stays in the arena
        st = eng.root_stats()
        torch.cuda.synchronize()
        assert eng.leaf_compaction and eng.leaf_compaction_fused == (fused == "1")
        outs.append({k: v.clone() for k, v in st.items()})
        search.close()
    for key in ("child_N", "child_W", "child_P", "root_W", "root_N", "legal"):
        assert torch.equal(outs[0][key], outs[1][key]), key
