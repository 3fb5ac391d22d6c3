"""Debug: where one warp of k_run_sims spends its cycles (library built with -DAZ_TRUNK_CLOCKS)."""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
import alphazero_implementation_b200 as az
from alphazero_implementation_b200 import _lib

E, S = 4096, 200
eng = az.Engine(num_games=E, num_simulations=S, device=0)
eng.reset_games()
u = torch.from_numpy(np.random.RandomState(0).random_sample((12, E))).cuda()
lib = _lib.load()
lib.az_debug_run_clocks.argtypes = [C.c_void_p, C.c_int]
for i in range(8):
    eng.run_simulations(S, 1)
    eng.sample_moves(u[i])
torch.cuda.synchronize()
lib.az_debug_run_clocks(None, 1)
eng.run_simulations(S, 1)
torch.cuda.synchronize()
buf = np.zeros(8, np.int64)
lib.az_debug_run_clocks(buf.ctypes.data, 0)
if len(sys.argv) > 1:  # fused move step: cycles of the move tail of the same warp
    eng.sample_moves(u[8])
    lib.az_debug_run_clocks(None, 1)
    t = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    t[0].record(); eng.run_simulations(S, 1); t[1].record(); eng.sample_moves(u[9]); torch.cuda.synchronize()
    t[1].record(); eng.run_move_step(S, 1, u[10]); t[2].record(); torch.cuda.synchronize()
    lib.az_debug_run_clocks(buf.ctypes.data, 0)
    print("move tail cycles: policy/draw/record %d  roots/ring header %d  episode copy %d ; fused kernel %.1f us, run_sims alone %.1f us" % (
        buf[4], buf[5], buf[6], t[1].elapsed_time(t[2]) * 1e3, 0.0))
print("per simulation (cycles): descend %.0f  leaf/expand %.0f  backup %.0f   total %.0f ; mean depth of lane-0 tree %.2f" % (
    buf[0] / S, buf[1] / S, buf[2] / S, (buf[0] + buf[1] + buf[2]) / S, buf[3] / S))
