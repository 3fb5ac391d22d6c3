import sys, torch
sys.path.insert(0, '.')
import alphazero_implementation_b200 as az
E = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
blocks = int(sys.argv[2]) if len(sys.argv) > 2 else 4
model = az.ResNet(blocks, 64)
s = az.AlphaZeroSearch(model=model, num_simulations=64, use_cuda_graph=False)
eng = s.engine_for(E); eng.reset_games()
for _ in range(8): s._graphed = None; 
from alphazero_implementation_b200.search import _GraphedStep
g = _GraphedStep(eng, s._net, s._net.input_layout); g.run(8, False)
net = s._net
eng.select_leaves()
def t(fn, n=20):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
h = net.trunk.forward_leaves(eng)
print("trunk us", t(lambda: net.trunk.forward_leaves(eng)))
print("policy head us", t(lambda: net.net.policy_head(h)))
print("value head us", t(lambda: net.net.value_head(h)))
print("full forward_leaves us", t(lambda: net.forward_leaves(eng)))
