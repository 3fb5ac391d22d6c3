"""GPU: the hand-written tcgen05 ResNet trunk (csrc/az_conv.cu) vs a plain PyTorch reference of the same op."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

import alphazero_implementation_b200 as az  # noqa: E402
from alphazero_implementation_b200.engine import LAYOUT_PLANES_F32  # noqa: E402
from alphazero_implementation_b200.models import TensorCoreTrunk, _fold_bn  # noqa: E402


def _engine_with_leaves(n, seed=0, deep=True, compact=False):
    """An engine whose last selection left a mix of leaves that wait for an evaluation and terminal leaves.  `compact`: the
    selection also writes the ordered list of the former (az_set_leaf_compaction), which the ResNet kernels then walk."""
    eng = az.Engine(num_games=n, num_simulations=48)
    eng.set_leaf_compaction(compact)
    eng.reset_games()
    u = torch.from_numpy(np.random.RandomState(seed).random_sample(n)).cuda()
    for _ in range(12 if deep else 3):
        eng.run_simulations(30, 2)
        eng.sample_moves(u)
    eng.run_simulations(30, 2)
    eng.select_leaves()
    return eng


def _randomise_bn(m):
    g = torch.Generator(device="cpu").manual_seed(1)
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.copy_(torch.randn(mod.num_features, generator=g) * 0.2)
            mod.running_var.copy_(torch.rand(mod.num_features, generator=g) + 0.5)
            mod.weight.data.copy_(torch.rand(mod.num_features, generator=g) + 0.5)
            mod.bias.data.copy_(torch.randn(mod.num_features, generator=g) * 0.2)


def _reference_trunk(model, x):
    """Same arithmetic in PyTorch: BN folded, bf16-rounded weights and inter-layer activations, fp32 accumulation."""
    r = lambda t: t.to(torch.bfloat16).to(torch.float32)
    w, b = _fold_bn(model.input_conv[0], model.input_conv[1])
    h = r(torch.relu(F.conv2d(r(x), r(w), b, padding=1)))
    for blk in model.residual_blocks:
        w1, b1 = _fold_bn(blk.conv1, blk.bn1)
        w2, b2 = _fold_bn(blk.conv2, blk.bn2)
        t = r(torch.relu(F.conv2d(h, r(w1), b1, padding=1)))
        h = r(torch.relu(F.conv2d(t, r(w2), b2, padding=1) + h))
    return h


@pytest.mark.parametrize("blocks,n", [(0, 8), (1, 5), (1, 64), (4, 100), (2, 1000), (1, 5003)])
def test_trunk_matches_pytorch_reference(blocks, n):
    torch.manual_seed(blocks * 100 + n)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    model = az.ResNet(num_res_blocks=blocks, num_channels=64).cuda().eval()
    _randomise_bn(model)
    eng = _engine_with_leaves(n, seed=n)
    status = eng.leaf_info()["status"]
    x = eng.gather_leaves(LAYOUT_PLANES_F32)
    trunk = TensorCoreTrunk(model, torch.device("cuda", torch.cuda.current_device()), variant=1)
    got = trunk.forward_leaves(eng).float()
    torch.cuda.synchronize()
    with torch.no_grad():
        ref = _reference_trunk(model, x)
    assert got.shape == ref.shape == (n, 64, 6, 7)
    assert torch.isfinite(got).all()
    live = status == 0
    assert live.any()
    err = (got[live] - ref[live]).abs()
    scale = ref[live].abs().max().item()
    # bf16 output rounding (2^-8 relative) + accumulation-order noise across 4..9 layers
    assert err.max().item() <= 0.03 * max(scale, 1.0), (err.max().item(), scale)
    assert err.mean().item() <= 2e-3 * max(scale, 1.0)
    eng.close()


def test_resnet_in_the_search_loop_matches_library_path():
    """ResNet 2x64 evaluated through the tcgen05 trunk (+ library heads) inside AlphaZeroSearch gives the same trees, up
    to bf16 accumulation-order noise, as the all-cuDNN bf16 path: root visit counts within a few visits."""
    torch.manual_seed(11)
    model = az.ResNet(num_res_blocks=2, num_channels=64)
    _randomise_bn(model)
    roots = [az.Config().sample_initial_state() for _ in range(4)]
    out = []
    for tc in (True, False):
        s = az.AlphaZeroSearch(model=model, num_simulations=96, use_tensor_core_kernels=tc, inference_dtype=torch.bfloat16)
        assert (s._net.trunk is not None) == tc
        nodes = [az.Node(r) for r in roots]
        s.run_simulations(nodes)
        out.append([[ch.visit_count for ch in n.children.values()] for n in nodes])
        assert all(sum(v) == 95 for v in out[-1])
    for a, b in zip(*out):
        assert max(abs(x - y) for x, y in zip(a, b)) <= 10


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("blocks,n", [(0, 8), (2, 77), (4, 1000), (2, 4999)])
def test_full_net_kernel_matches_pytorch_reference(blocks, n, variant):
    """Trunk + fused heads (policy conv1x1 + FC, value conv3x3 + FC + tanh) vs the fp32 PyTorch module evaluated on the
    bf16-emulated trunk: logits / values within bf16 noise, and vs the plain fp32 module within the bf16 budget."""
    torch.manual_seed(7 * blocks + n)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    model = az.ResNet(num_res_blocks=blocks, num_channels=64).cuda().eval()
    _randomise_bn(model)
    eng = _engine_with_leaves(n, seed=n + 1, compact=n % 2 == 1)  # both ways of walking the batch
    live = eng.leaf_info()["status"] == 0
    x = eng.gather_leaves(LAYOUT_PLANES_F32)
    trunk = TensorCoreTrunk(model, torch.device("cuda", torch.cuda.current_device()), variant=variant)  # both 64-channel kernels
    logits, values = trunk.forward_leaves_full(eng)
    torch.cuda.synchronize()
    r = lambda t: t.to(torch.bfloat16).to(torch.float32)
    with torch.no_grad():
        h = _reference_trunk(model, x)
        wp, bp = _fold_bn(model.policy_head[0], model.policy_head[1])
        wv, bv = _fold_bn(model.value_head[0], model.value_head[1])
        pa = torch.relu(F.conv2d(h, r(wp), bp))
        va = torch.relu(F.conv2d(h, r(wv), bv, padding=1))
        l_ref = model.policy_head[4](pa.flatten(1))
        v_ref = torch.tanh(model.value_head[4](va.flatten(1)))
        l_fp32, v_fp32 = model(x)
    assert torch.isfinite(logits).all() and torch.isfinite(values).all()
    assert torch.allclose(logits[live], l_ref[live], atol=3e-2, rtol=2e-2), float((logits[live] - l_ref[live]).abs().max())
    assert torch.allclose(values[live, :1], v_ref[live], atol=2e-2), float((values[live, :1] - v_ref[live]).abs().max())
    assert torch.equal(values[:, 1], -values[:, 0])
    assert float((torch.softmax(logits[live], 1) - torch.softmax(l_fp32[live], 1)).abs().max()) < 5e-2
    assert float((values[live] - v_fp32[live]).abs().max()) < 8e-2
    eng.close()


@pytest.mark.parametrize("blocks,n", [(1, 5), (3, 1001), (1, 3333)])
def test_cta_pair_variant_is_bit_identical(blocks, n):
    """The cta_group::2 variant of the kernel (two CTAs per MMA, M = 256) must give exactly the single-CTA results (which for
    more than 148 x 8 positions are produced by persistent CTAs walking over several batches),
    including an odd tail (the last pair has an empty partner)."""
    from alphazero_implementation_b200 import _lib
    from alphazero_implementation_b200.models import InferenceNet

    lib = _lib.load()
    torch.manual_seed(7 + n)
    model = az.ResNet(num_res_blocks=blocks, num_channels=64).cuda().eval()
    _randomise_bn(model)
    eng = _engine_with_leaves(n, seed=n)
    net = InferenceNet(model, dtype=torch.bfloat16, trunk_variant=1)
    assert net.evaluates_leaves_directly and net.kernel_name == "k_resnet_trunk"
    outs = []
    try:
        for pair in (0, 1):
            lib.az_trunk_set_cta_pair(pair)
            trunk_out = net.trunk.forward_leaves(eng).clone()
            logits, values = net.forward_leaves(eng)
            outs.append((trunk_out, logits.clone(), values.clone()))
        torch.cuda.synchronize()
    finally:
        lib.az_trunk_set_cta_pair(0)
    for a, b in zip(*outs):
        assert torch.equal(a, b)
    eng.close()
