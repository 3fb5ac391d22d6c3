"""GPU: the hand-written tcgen05 kernels for the reference's CNNModel (csrc/az_cnn.cu; models/games/connect4/cnn.py:8-75) against
plain PyTorch: (a) the same arithmetic emulated (BatchNorm folded, 16-bit-rounded weights and activations between layers, fp32
accumulation), (b) the fp32 module within north_star's 1e-3 on priors / values in fp16 mode."""
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
    h = r(x)
    for i in (0, 3, 6):
        w, b = _fold_bn(model.conv_layers[i], model.conv_layers[i + 1])
        h = r(torch.relu(F.conv2d(h, r(w), b, padding=1)))
    fc = model.shared_layers[0]
    s = torch.relu(F.linear(h.flatten(1), r(fc.weight), fc.bias))  # the hidden layer stays fp32 in the kernel (heads on CUDA cores)
    return model.policy_head(s), model.value_head(s)


@pytest.mark.parametrize("n,dtype,compact", [(1, torch.bfloat16, True), (5, torch.float16, True), (128, torch.bfloat16, False), (129, torch.float16, True),
                                             (1000, torch.bfloat16, True), (2501, torch.float16, True), (4100, torch.bfloat16, False)])
def test_cnn_kernels_match_pytorch(n, dtype, compact):
    torch.manual_seed(100 + n)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    model = az.CNNModel().cuda().eval()
    _randomise_bn(model)
    eng = _engine_with_leaves(n, seed=n + 5, compact=compact)
    live = eng.leaf_info()["status"] == 0
    assert live.any()
    x = eng.gather_leaves(LAYOUT_PLANES_F32)
    net = InferenceNet(model, dtype=dtype)
    assert net.kernel_name == "k_cnn_conv + k_cnn_fc" and net.evaluates_leaves_directly
    logits, values = net.forward_leaves(eng)
    torch.cuda.synchronize()
    with torch.no_grad():
        l_emu, v_emu = _emulated(model, x, dtype)
        l_32, v_32 = model(x)
    assert torch.isfinite(logits[live]).all() and torch.isfinite(values[live]).all()
    if compact:
        assert (logits[~live] == 0).all() and (values[~live] == 0).all()
    tol = 3e-3 if dtype == torch.float16 else 2e-2
    assert torch.allclose(logits[live], l_emu[live], atol=tol, rtol=2e-2), float((logits[live] - l_emu[live]).abs().max())
    assert torch.allclose(values[live, :1], v_emu[live], atol=tol), float((values[live, :1] - v_emu[live]).abs().max())
    assert torch.equal(values[live, 1], -values[live, 0])
    dp = float((torch.softmax(logits[live], 1) - torch.softmax(l_32[live], 1)).abs().max())
    dv = float((values[live] - v_32[live]).abs().max())
    if dtype == torch.float16:
        assert dp <= 1e-3 and dv <= 1e-3, (dp, dv)  # north_star tolerance against the fp32 `predict`
    else:
        assert dp <= 1e-2 and dv <= 3e-2, (dp, dv)
    eng.close()


def test_cnn_in_the_search_loop():
    torch.manual_seed(9)
    model = az.CNNModel()
    search = az.AlphaZeroSearch(model=model, num_simulations=48, inference_dtype=torch.bfloat16)
    assert search.evaluator_name == "k_cnn_conv + k_cnn_fc"
    nodes = [az.Node(az.Config().sample_initial_state()) for _ in range(5)]
    search.run_simulations(nodes)
    a = [[ch.visit_count for ch in nd.children.values()] for nd in nodes]
    assert all(sum(v) == 47 for v in a) and all(v == a[0] for v in a)
    # against the library path (cuDNN / cuBLAS bf16): same trees up to rounding noise
    lib = az.AlphaZeroSearch(model=model, num_simulations=48, inference_dtype=torch.bfloat16, use_tensor_core_kernels=False)
    nodes = [az.Node(az.Config().sample_initial_state()) for _ in range(5)]
    lib.run_simulations(nodes)
    b = [[ch.visit_count for ch in nd.children.values()] for nd in nodes]
    assert max(abs(x - y) for x, y in zip(a[0], b[0])) <= 8
    search.close()
    lib.close()
