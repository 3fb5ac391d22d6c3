"""GPU: the layer-pipelined tcgen05 ResNet kernel (csrc/az_resnet_pipe.cu) - 128 channels, the reference's own
ResNet(board, 7, 9, 128) (src/alphazero_simple/resnet.py:30-103, src/alphazero_less_simple/main.py:13), and 64 channels, the
headline net - against plain PyTorch:
(a) the same arithmetic emulated in PyTorch (BatchNorm folded, 16-bit-rounded weights and inter-layer activations, fp32
accumulation) within accumulation-order noise, (b) the fp32 module within north_star's 1e-3 on priors / values in fp16 mode."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

import alphazero_implementation_b200 as az  # noqa: E402
from alphazero_implementation_b200.engine import LAYOUT_PLANES_F32  # noqa: E402
from alphazero_implementation_b200.models import InferenceNet, _fold_bn  # noqa: E402
from test_gpu_trunk import _engine_with_leaves, _randomise_bn  # noqa: E402


def _emulated(model, x, dtype):
    r = lambda t: t.to(dtype).to(torch.float32)
    w, b = _fold_bn(model.input_conv[0], model.input_conv[1])
    h = r(torch.relu(F.conv2d(r(x), r(w), b, padding=1)))
    for blk in model.residual_blocks:
        w1, b1 = _fold_bn(blk.conv1, blk.bn1)
        w2, b2 = _fold_bn(blk.conv2, blk.bn2)
        t = r(torch.relu(F.conv2d(h, r(w1), b1, padding=1)))
        h = r(torch.relu(F.conv2d(t, r(w2), b2, padding=1) + h))
    wp, bp = _fold_bn(model.policy_head[0], model.policy_head[1])
    wv, bv = _fold_bn(model.value_head[0], model.value_head[1])
    pa = torch.relu(F.conv2d(h, r(wp), bp))
    va = torch.relu(F.conv2d(h, r(wv), bv, padding=1))
    return model.policy_head[4](pa.flatten(1)), torch.tanh(model.value_head[4](va.flatten(1)))


@pytest.mark.parametrize("channels,blocks,n,dtype", [
    (128, 0, 4, torch.bfloat16), (128, 1, 3, torch.bfloat16), (128, 1, 64, torch.float16), (128, 2, 1000, torch.bfloat16),
    (128, 9, 700, torch.float16), (128, 3, 2501, torch.bfloat16),
    (64, 0, 8, torch.bfloat16), (64, 1, 5, torch.float16), (64, 4, 64, torch.float16), (64, 4, 3001, torch.bfloat16), (64, 11, 1190, torch.float16),
    (-64, 0, 4, torch.bfloat16), (-64, 1, 7, torch.float16), (-64, 4, 64, torch.float16), (-64, 4, 3001, torch.bfloat16), (-64, 5, 2381, torch.float16),
    (-640, 0, 4, torch.bfloat16), (-640, 1, 7, torch.float16), (-640, 1, 8, torch.float16), (-640, 4, 64, torch.float16), (-640, 4, 3001, torch.bfloat16),
    (-640, 5, 2381, torch.float16),
    (-1280, 0, 4, torch.bfloat16), (-1280, 1, 5, torch.float16), (-1280, 2, 1000, torch.bfloat16), (-1280, 9, 701, torch.float16),
    (-192, 0, 4, torch.bfloat16), (-192, 0, 9, torch.float16), (-192, 1, 7, torch.float16), (-192, 1, 8, torch.bfloat16), (-192, 4, 64, torch.float16),
    (-192, 4, 3001, torch.bfloat16), (-192, 4, 2381, torch.float16), (-192, 11, 1190, torch.float16), (-192, 2, 593, torch.float16)])
def test_resnet_pipe_kernel_matches_pytorch(channels, blocks, n, dtype):
    # -64: two CTAs per SM; -640 / -1280: CTA pairs (64 / 128 channels); -192: filter rows fused, N = 192 (csrc/az_resnet_wide.cu)
    variant = {64: 0, 128: 0, -64: 2, -640: 3, -1280: 3, -192: 4}[channels]
    channels = 128 if channels in (128, -1280) else 64
    torch.manual_seed(13 * blocks + n)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    model = az.ResNet(num_res_blocks=blocks, num_channels=channels).cuda().eval()
    _randomise_bn(model)
    compact = n != 64  # one case computes every slot's row, the others walk the compacted leaf list
    eng = _engine_with_leaves(n, seed=n + 2, compact=compact)
    live = eng.leaf_info()["status"] == 0
    assert live.any()
    x = eng.gather_leaves(LAYOUT_PLANES_F32)
    net = InferenceNet(model, dtype=dtype, trunk_variant=variant)
    assert net.kernel_name == ("k_resnet_wide" if variant == 4 else "k_resnet_pipe")
    logits, values = net.forward_leaves(eng)
    torch.cuda.synchronize()
    with torch.no_grad():
        l_emu, v_emu = _emulated(model, x, dtype)
        l_32, v_32 = model(x)
    assert torch.isfinite(logits).all() and torch.isfinite(values).all()
    if compact:
        assert (logits[~live] == 0).all() and (values[~live] == 0).all()  # slots without an evaluation are left alone
    tol = 4e-3 if dtype == torch.float16 else 3e-2  # accumulation order differs (taps outer vs K chunks outer); one rounding flip per activation
    assert torch.allclose(logits[live], l_emu[live], atol=tol, rtol=2e-2), float((logits[live] - l_emu[live]).abs().max())
    assert torch.allclose(values[live, :1], v_emu[live], atol=tol), float((values[live, :1] - v_emu[live]).abs().max())
    assert torch.equal(values[:, 1], -values[:, 0])
    dp = float((torch.softmax(logits[live], 1) - torch.softmax(l_32[live], 1)).abs().max())
    dv = float((values[live] - v_32[live]).abs().max())
    if dtype == torch.float16:
        assert dp <= 1e-3 and dv <= 1e-3, (dp, dv)  # north_star tolerance against the fp32 `predict`
    else:
        assert dp <= 2e-2 and dv <= 5e-2, (dp, dv)
    eng.close()


def test_resnet128_in_the_search_loop_and_weight_refresh():
    """9 x 128 through AlphaZeroSearch (CUDA-graphed steps); `update_inference_model` rewrites the packed weights in place and the
    captured graph keeps working."""
    torch.manual_seed(3)
    model = az.ResNet(num_res_blocks=9, num_channels=128)
    search = az.AlphaZeroSearch(model=model, num_simulations=40, inference_dtype=torch.bfloat16)
    assert search.evaluator_name == "k_resnet_pipe"
    nodes = [az.Node(az.Config().sample_initial_state()) for _ in range(6)]
    search.run_simulations(nodes)
    a = [[ch.visit_count for ch in nd.children.values()] for nd in nodes]
    assert all(sum(v) == 39 for v in a) and all(v == a[0] for v in a)  # same root, same evaluator: same tree
    graph = search._graphed.graph
    model2 = az.ResNet(num_res_blocks=9, num_channels=128)
    search.update_inference_model(model2)
    assert search._graphed is not None and search._graphed.graph is graph  # refreshed in place
    nodes = [az.Node(az.Config().sample_initial_state()) for _ in range(6)]
    search.run_simulations(nodes)
    b = [[ch.visit_count for ch in nd.children.values()] for nd in nodes]
    fresh = az.AlphaZeroSearch(model=model2, num_simulations=40, inference_dtype=torch.bfloat16)
    nodes = [az.Node(az.Config().sample_initial_state()) for _ in range(6)]
    fresh.run_simulations(nodes)
    assert b == [[ch.visit_count for ch in nd.children.values()] for nd in nodes]
    search.close()
    fresh.close()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_wide_kernel_is_deterministic_and_independent_of_the_batch_composition(dtype):
    """k_resnet_wide hands tiles back and forth between the tensor core, eight epilogue warps, a stager warp and an FC warp through
    mbarriers: (a) two launches on the same leaves give bit-identical outputs (no race decides anything); (b) a position's outputs do
    not depend on where it sits - with leaf compaction (position j = the j-th leaf that waits for an evaluation) and without (position
    = slot, terminal leaves as empty rows) every evaluated slot gets the same bits, because its neighbours in a tile only ever
    contribute exact zeros (the shared zero row / column of the compact padding)."""
    torch.manual_seed(5)
    model = az.ResNet(num_res_blocks=3, num_channels=64).cuda().eval()
    _randomise_bn(model)
    outs = []
    for compact in (True, False):
        eng = _engine_with_leaves(2999, seed=11, compact=compact)
        live = eng.leaf_info()["status"] == 0
        net = InferenceNet(model, dtype=dtype, trunk_variant=4)
        assert net.kernel_name == "k_resnet_wide"
        l1, v1 = (t.clone() for t in net.forward_leaves(eng))
        l2, v2 = (t.clone() for t in net.forward_leaves(eng))
        torch.cuda.synchronize()
        assert torch.equal(l1, l2) and torch.equal(v1, v2)
        outs.append((l1[live], v1[live]))
        eng.close()
    assert outs[0][0].shape[0] > 2000
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


def test_wide_kernel_timing_hook_counts_launches_and_stamps_durations():
    """`az_resnet_wide_set_timing` (what bench.py times the headline kernel with inside its timed region): the device buffer counts the
    launches and holds {first CTA start, last CTA end} of each on the global timer; switching the hook off stops the counting."""
    from alphazero_implementation_b200 import _lib

    lib = _lib.load()
    cap = 8
    buf = torch.zeros(4 + 2 * cap, dtype=torch.int64, device="cuda")
    buf[4::2] = -1
    torch.manual_seed(1)
    model = az.ResNet(num_res_blocks=2, num_channels=64).cuda().eval()
    eng = _engine_with_leaves(500, seed=3, compact=True)
    net = InferenceNet(model, dtype=torch.bfloat16, trunk_variant=4)
    try:
        assert lib.az_resnet_wide_set_timing(buf.data_ptr(), cap) == 0
        for _ in range(3):
            net.forward_leaves(eng)
        torch.cuda.synchronize()
        assert int(buf[0]) == 3 and int(buf[1]) == 0
        dur = (buf[5:5 + 6:2] - buf[4:4 + 6:2]).cpu()
        assert (dur > 1_000).all() and (dur < 50_000_000).all(), dur  # ns: a launch of 500 positions takes tens of microseconds
        assert int(buf[4 + 2 * 3]) == -1 and int(buf[5 + 2 * 3]) == 0  # untouched slot
    finally:
        lib.az_resnet_wide_set_timing(None, 0)
    net.forward_leaves(eng)
    torch.cuda.synchronize()
    assert int(buf[0]) == 3
    eng.close()
