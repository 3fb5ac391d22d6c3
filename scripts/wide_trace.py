"""Timing study: timeline of CTA 0 of k_resnet_wide (library built with -DWIDE_TRACE; AZ_ENGINE_LIB points at it)."""
import ctypes as C
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
import alphazero_implementation_b200 as az
from alphazero_implementation_b200 import _lib

blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 2
E = 16384
search = az.AlphaZeroSearch(model=az.ResNet(blocks, 64), num_simulations=64, use_cuda_graph=False, trunk_variant=4)
eng = search.engine_for(E)
eng.reset_games()
net = search._net
lib = _lib.load()
lib.az_resnet_wide_trace.restype = C.c_int32
lib.az_resnet_wide_trace.argtypes = [C.c_void_p, C.c_int32]
buf = np.zeros(3 * 340 * 2, np.uint32)
for it in range(3):
    eng.select_leaves()
    logits, values = net.forward_leaves(eng)
    eng.expand_backup(logits, values, 1)
    torch.cuda.synchronize()
assert lib.az_resnet_wide_trace(buf.ctypes.data, buf.size) == buf.size
ev = buf.reshape(3, 340, 2)
names = ["issuer", "epi-w0", "epi-w7"]
rows = [(int(c), names[w], int(code)) for w in range(3) for code, c in ev[w] if code != 0xFFFFFFFF]
t0 = min(r[0] for r in rows)
rows = sorted(((c - t0) & 0xFFFFFFFF, who, code) for c, who, code in rows)
for t, who, code in rows[:int(sys.argv[2]) if len(sys.argv) > 2 else 600]:
    print(f"{t:8d}  {who:7s} layer {code // 100:2d} tile {code // 10 % 10} phase {code % 10}")
