"""Debug: per-layer timestamps of one CTA of the tcgen05 ResNet kernel (library built with -DAZ_TRUNK_CLOCKS)."""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
import alphazero_implementation_b200 as az
from alphazero_implementation_b200 import _lib

E = 16384
model = az.ResNet(4, 64)
s = az.AlphaZeroSearch(model=model, num_simulations=64, use_cuda_graph=False)
eng = s.engine_for(E)
eng.reset_games()
eng.select_leaves()
net = s._net
for _ in range(3):
    net.forward_leaves(eng)
torch.cuda.synchronize()
lib = _lib.load()
buf = np.zeros(4 * 2 * 24, np.int64)
lib.az_debug_trunk_clocks.argtypes = [C.c_void_p]
assert lib.az_debug_trunk_clocks(buf.ctypes.data) == 0
k = buf.reshape(4, 24, 2)
t0 = k[0, 23, 0]
print("kernel CTA total cycles", k[1, 23, 0] - t0)
for l in range(10):
    for g in range(2):
        if k[0, l, g] == 0:
            continue
        print(f"layer {l} g{g}: issue_start {k[0,l,g]-t0:7d} issue_end {k[1,l,g]-t0:7d} acc_ready {k[2,l,g]-t0:7d} epi_end {k[3,l,g]-t0:7d}"
              f" | issue {k[1,l,g]-k[0,l,g]:6d} mma(after issue start) {k[2,l,g]-k[0,l,g]:6d} epi {k[3,l,g]-k[2,l,g]:6d}")
