"""Where does a network-in-the-loop simulation step spend its time?  (select / gather / net / expand+backup)"""
import sys
import torch
sys.path.insert(0, '.')
import alphazero_implementation_b200 as az
from alphazero_implementation_b200.engine import POLICY_LOGITS

E = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
nets = sys.argv[2:] or ["resnet4x64", "resnet2x64", "resnet9x128", "basic", "cnn"]
torch.backends.cudnn.benchmark = True
for name in nets:
    kw = {}
    if name == "basic_tc":  # BasicNN through the hand-written tcgen05 kernel
        model = az.BasicNN()
        kw = dict(inference_dtype=torch.bfloat16)
    elif name == "basic":
        model = az.BasicNN()
    elif name.endswith("_lib"):  # ResNet through cuDNN only (no hand-written trunk)
        b, c = name[:-4].replace("resnet", "").split("x")
        model = az.ResNet(int(b), int(c))
        kw = dict(use_tensor_core_kernels=False)
    elif name == "cnn":
        model = az.CNNModel()
    else:
        base, _, v = name.partition(":v")  # "resnet4x64:v4" = trunk_variant 4
        b, c = base.replace("resnet", "").split("x")
        model = az.ResNet(int(b), int(c))
        if v:
            kw = dict(trunk_variant=int(v))
    search = az.AlphaZeroSearch(model=model, num_simulations=64, use_cuda_graph=False, **kw)
    eng = search.engine_for(E)
    eng.reset_games()
    net = search._net
    x = eng.gather_leaves(net.input_layout)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    tot = [0.0] * 4
    for it in range(40):
        ev[0].record(); eng.select_leaves()
        ev[1].record()
        direct = net.evaluates_leaves_directly
        if not direct:
            eng.gather_leaves(net.input_layout, x)
        ev[2].record()
        logits, values = net.forward_leaves(eng) if direct else net(x)
        ev[3].record(); eng.expand_backup(logits, values, POLICY_LOGITS)
        ev[4].record()
        torch.cuda.synchronize()
        if it >= 8:
            for k in range(4):
                tot[k] += ev[k].elapsed_time(ev[k + 1])
    n = 32
    print(f"{name:12s} E={E}: select {tot[0]/n*1e3:7.1f} us  gather {tot[1]/n*1e3:7.1f} us  net {tot[2]/n*1e3:8.1f} us  expand {tot[3]/n*1e3:7.1f} us"
          f"  -> {E/(sum(tot)/n)*1e3:.3e} sims/s eager", flush=True)
    eng.close()
