"""Debug: phase timestamps of one CTA of k_mlp_fused (library built with -DAZ_TRUNK_CLOCKS)."""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
import alphazero_implementation_b200 as az
from alphazero_implementation_b200 import _lib

E = 16384
s = az.AlphaZeroSearch(model=az.BasicNN(), num_simulations=64, use_cuda_graph=False, inference_dtype=torch.bfloat16)
eng = s.engine_for(E)
eng.reset_games()
eng.select_leaves()
net = s._net
for _ in range(3):
    net.forward_leaves(eng)
torch.cuda.synchronize()
lib = _lib.load()
buf = np.zeros(16, np.int64)
lib.az_debug_mlp_clocks.argtypes = [C.c_void_p]
assert lib.az_debug_mlp_clocks(buf.ctypes.data) == 0
names = ["start", "L1 issue start", "L1 issued(thread0)", "L1 acc ready", "L1 epilogue end", "L2 issue start", "L2 issued", "L2 acc ready",
         "L2 epilogue end", "head issue start", "head acc ready", "end"]
for i, nm in enumerate(names):
    print(f"{nm:22s} {buf[i]-buf[0]:8d}  (+{buf[i]-buf[max(i-1,0)]:6d})")
